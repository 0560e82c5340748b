"""Quick device-time measurement of one BER configuration (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wofdm_b200 as W
from wofdm_b200 import capi

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "wtx"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    ens = int(sys.argv[3]) if len(sys.argv) > 3 else 27
    scale = N // 256
    cp, ttx, trx = 16 * scale, 8 * scale, 10 * scale
    if name in ("CP", "wrx", "CPwrx"): ttx = 0
    if name in ("CP", "wtx", "CPwtx"): trx = 0
    s = W.params_from_name(name, N, cp, ttx, trx, bits=4 if N == 256 else 6, S=16)
    h = W.Handle([0])
    print("fp32 peak scalar", h.fp32_peak(0), "ffma2", h.fp32_peak(1), "cmac pattern", h.fp32_peak(2), "ffma2+lop3 1:1", h.fp32_peak(3))
    vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
    rng = np.random.default_rng(0)
    C = 250
    chan = (rng.standard_normal((21, C)) + 1j * rng.standard_normal((21, C))) * np.exp(-np.arange(21) / 4)[:, None]
    snr = np.linspace(-20, 50, 30)
    plan = h.ber_plan(s, vt, vr, chan, snr)
    print("kernel", plan.kernel)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    assert st != 0
    for _ in range(3):
        plan.launch(ens, seed=1, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    e0.record()
    for k in range(K):
        plan.launch(ens, seed=2 + k, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    frames = 30 * C * ens
    syms = frames * 16
    print(f"{name} N={N}: {ms:.3f} ms/launch, {frames} frames, {syms / ms * 1e3:.4g} OFDM symbols/s")
    be, se = plan.read()
    bt, stot = plan.totals(ens)
    print("SER", np.round(se / stot, 4)[::3])
    t0 = time.time()
    r = h.ber_run(s, vt, vr, chan, snr, ens, seed=3)
    print("e2e ber_run", time.time() - t0, "s")

main()
