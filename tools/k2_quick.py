"""Wall / device times of wofdm_interf_power per mode (development aid): k2_quick.py [C] [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi
C = int(sys.argv[1]) if len(sys.argv) > 1 else 250
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sc = N // 256
s = W.params_from_name("WOLA", N, 16 * sc, 8 * sc, 10 * sc, bits=4, S=16)
vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
rng = np.random.default_rng(0)
chan = (rng.standard_normal((21, C)) + 1j * rng.standard_normal((21, C))) * np.exp(-np.arange(21) / 4)[:, None]
with W.Handle([0]) as h:
    ref = None
    for mode in (0, 1, 2):
        if mode == 1 and N % 256: continue
        for _ in range(3): P = h.interf_power(s, vt, vr, chan, mode=mode)
        t0 = time.perf_counter(); K = 10
        for _ in range(K): P = h.interf_power(s, vt, vr, chan, mode=mode)
        wall = (time.perf_counter() - t0) / K * 1e3
        tm = h.interf_last_timing()
        if ref is None: ref = P
        print(f"mode {mode}: C={C} N={N} wall {wall:.3f} ms, device {tm}, max rel diff vs mode 0 {np.max(np.abs(P - ref)) / np.max(np.abs(ref)):.2e}", flush=True)
