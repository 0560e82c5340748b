"""Throughput of the channel-mask BER variant (matlab/main_channel_mask.m): masked chain (mask product on the tensor cores,
mask_gemm.cu, + the K1 kernel that gathers from its output; WOFDM_MASK_FFT=1: the per-symbol FFT kernel) and the unmasked
guard-band chain on the same symbols; wall time of the C-ABI calls with host buffers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi

h = W.Handle([0])
s = W.params_from_name("wtx", 256, 16, 8, 0, bits=4, S=16, noise_norm=1, constellation=1, guard=64)
vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
rng = np.random.default_rng(0)
chan = (rng.standard_normal((21, 250)) + 1j * rng.standard_normal((21, 250))) * np.exp(-np.arange(21) / 4)[:, None]
snr = np.linspace(-20, 50, 30)
for ens in (1, 4):
    h.ber_run_masked(s, vt, vr, chan, snr, ens, seed=1, variant=1)
    t0 = time.perf_counter(); m = h.ber_run_masked(s, vt, vr, chan, snr, ens, seed=1, variant=1); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); u = h.ber_run(s, vt, vr, chan, snr, ens, seed=1, variant=0); t2 = time.perf_counter() - t0
    syms = 30 * 250 * ens * 16
    print(f"ensemble {ens}: masked {t1*1e3:.1f} ms = {syms/t1:.3g} OFDM symbols/s | unmasked guard-band {t2*1e3:.2f} ms = {syms/t2:.3g} symbols/s")
    print("   BER masked  ", np.round(m["bit_err"] / m["bit_tot"], 4)[::6])
    print("   BER unmasked", np.round(u["bit_err"] / u["bit_tot"], 4)[::6])
