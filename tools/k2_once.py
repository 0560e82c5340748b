"""One wofdm_interf_power call per mode on the configs[3] shapes (for ncu launch lists): k2_once.py [C]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi
C = int(sys.argv[1]) if len(sys.argv) > 1 else 250
s = W.params_from_name("WOLA", 256, 16, 8, 10, precision=1)
vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
rng = np.random.default_rng(0)
chan = (rng.standard_normal((21, C)) + 1j * rng.standard_normal((21, C))) * np.exp(-np.arange(21) / 4)[:, None]
with W.Handle([0]) as h:
    for mode in (0, 1, 2):
        P = h.interf_power(s, vt, vr, chan, mode=mode)
        print("mode", mode, float(P.sum()))
