"""K2 timing across batch budgets (development aid): python tools/k2_ab.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi
s = W.params_from_name("WOLA", 256, 16, 8, 10, precision=1)
rng = np.random.default_rng(0)
C = 250
chan = (rng.standard_normal((21, C)) + 1j * rng.standard_normal((21, C))) * np.exp(-np.arange(21) / 4)[:, None]
h = W.Handle([0])
vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
ref = None
for mb in sys.argv[1:] or ["1024", "96", "64", "48", "32", "24", "16"]:
    os.environ["WOFDM_K2_BATCH_MB"] = mb
    for mode in (0, 1):
        for _ in range(3):
            P = h.interf_power(s, vt, vr, chan, mode=mode)
        t0 = time.perf_counter()
        for _ in range(10):
            P = h.interf_power(s, vt, vr, chan, mode=mode)
        dt = (time.perf_counter() - t0) / 10
        if ref is None: ref = P
        print(f"batch {mb:>5s} MB mode {mode}: {dt*1e3:.3f} ms  rel diff {np.abs(P-ref).max()/np.abs(ref).max():.2e}", flush=True)
