"""One masked call (wtx N=256, 250 channels x 30 SNR points x ensemble) for kernel launch lists (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi

ens = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h = W.Handle([0])
rng = np.random.default_rng(0)
snr = np.linspace(-20, 50, 30)
s = W.params_from_name("wtx", 256, 16, 8, 0, bits=4, S=16, noise_norm=1, constellation=1, guard=64)
vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
chan = (rng.standard_normal((21, 250)) + 1j * rng.standard_normal((21, 250))) * np.exp(-np.arange(21) / 4)[:, None]
for k in range(3):
    t0 = time.perf_counter(); h.ber_run_masked(s, vt, vr, chan, snr, ens, seed=1, variant=1); t1 = time.perf_counter() - t0
    print(f"call {k}: {t1 * 1e3:.2f} ms = {30 * 250 * ens * 16 / t1:.3g} OFDM symbols/s", flush=True)
