"""Channel-mask Tx side, dense tensor-core product (mask_gemm.cu) against the per-symbol FFT kernel (mask_kernel.cuh):
masked Tx streams of the first frames (WOFDM_MASK_DUMP), counters and wall time of the whole masked call (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi

h = W.Handle([0])
rng = np.random.default_rng(0)
snr = np.linspace(-20, 50, 30)
for name, N, cp, ttx, trx, bits, guard in (("wtx", 256, 16, 8, 0, 4, 64), ("WOLA", 256, 22, 8, 10, 6, 64), ("CP", 128, 8, 0, 0, 2, 0), ("CPW", 512, 32, 16, 20, 4, 128)):
    s = W.params_from_name(name, N, cp, ttx, trx, bits=bits, S=16, noise_norm=1, constellation=1, guard=guard)
    vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
    chan = (rng.standard_normal((21, 250)) + 1j * rng.standard_normal((21, 250))) * np.exp(-np.arange(21) / 4)[:, None]
    out = {}
    for mode in ("1", "0"):
        os.environ["WOFDM_MASK_FFT"] = mode
        os.environ["WOFDM_MASK_DUMP"] = f"/tmp/mask_{mode}.bin"
        r = h.ber_run_masked(s, vt, vr, chan, snr, 1, seed=1, variant=1)
        del os.environ["WOFDM_MASK_DUMP"]
        for ens in (1, 4):
            h.ber_run_masked(s, vt, vr, chan, snr, ens, seed=1, variant=1)
            t0 = time.perf_counter(); h.ber_run_masked(s, vt, vr, chan, snr, ens, seed=1, variant=1); t1 = time.perf_counter() - t0
            print(f"{name} N={N} {'fft ' if mode == '1' else 'gemm'} ensemble {ens}: {t1 * 1e3:.2f} ms = {30 * 250 * ens * 16 / t1:.3g} OFDM symbols/s", flush=True)
        out[mode] = (np.fromfile(f"/tmp/mask_{mode}.bin", dtype=np.float32), r)
    a, b = out["1"][0], out["0"][0]
    print(f"{name} N={N}: stream max |fft| {np.abs(a).max():.4g}, max |gemm - fft| {np.abs(a - b).max():.3g}, rms {np.sqrt(np.mean((a - b) ** 2)):.3g}; "
          f"counter differences bit {np.abs(out['1'][1]['bit_err'] - out['0'][1]['bit_err']).max()} sym {np.abs(out['1'][1]['sym_err'] - out['0'][1]['sym_err']).max()}", flush=True)
