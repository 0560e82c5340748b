"""Throughput of the PSD / out-of-band-radiation estimate (next row 8f-3): wall time of wofdm_psd_estimate with host
buffers for 1 and many records of 256 OFDM symbols (N = 256, 2048-point periodogram)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi

h = W.Handle([0])
s = W.params_from_name("WOLA", 256, 16, 8, 10)
w = capi.rc_window_tx(s)
for records in (1, 64, 4096):
    h.psd_estimate(256, 16, s.cs, 8, w, records=records, seed=1)
    t0 = time.perf_counter()
    x = h.psd_estimate(256, 16, s.cs, 8, w, records=records, seed=2)
    dt = time.perf_counter() - t0
    print(f"records {records}: {dt*1e3:.2f} ms = {records * 256 / dt:.3g} OFDM symbols/s, {records / dt:.3g} records/s; OBR {np.mean(np.r_[x[:384], x[-384:]]):.3e}")
