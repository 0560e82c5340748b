"""Interference-power benchmark (BASELINE.json configs[3]): per-channel P for C channels, one (system, window) pair.
Reports the dense-contraction rate (8*N*n_rx*n_tx + 8*N*n_tx*N flop per (channel, slice)) and the rate of
the tensor-core GEMM actually executed (8*N*n_rx*N per slice)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "WOLA"
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ttx = 8 if name in ("CPW", "WOLA", "CPwtx", "wtx") else 0
    trx = 10 if name in ("CPW", "WOLA", "CPwrx", "wrx") else 0
    s = W.params_from_name(name, 256, 16, ttx, trx, precision=1)
    rng = np.random.default_rng(0)
    chan = (rng.standard_normal((21, C)) + 1j * rng.standard_normal((21, C))) * np.exp(-np.arange(21) / 4)[:, None]
    h = W.Handle([0])
    vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
    P = h.interf_power(s, vt, vr, chan, mode=mode)      # warm-up (arena allocation, module load)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        P = h.interf_power(s, vt, vr, chan, mode=mode)
    dt = (time.perf_counter() - t0) / reps
    n_rx, n_tx, N = s.stride, s.n_tx, s.N
    M = 1 + -(-(21 - 1 + ttx) // n_rx)
    dense = (8 * N * n_rx * n_tx + 8 * N * n_tx * N) * C * M
    gemm = 8 * N * n_rx * N * C * M
    print(f"{name} C={C} mode={mode}: {dt*1e3:.2f} ms per call (host buffers in/out), dense-contraction rate "
          f"{dense/dt/1e12:.2f} TFLOP/s, executed GEMM {gemm/dt/1e12:.2f} TFLOP/s, P total {P.sum():.6g}")

main()
