"""Driver-level interference sweep (BASELINE.json configs[3]): every (system, window variant) x CP length x channel
realisation.  22 (system, variant) pairs per CP (wtx 2, CPwtx 2, wrx 2, CPwrx 2, WOLA 7, CPW 7) + CP-OFDM, 12 CP
lengths, C = 250 channels, fp64 DMMA (mode 0) and TF32-split tcgen05 (mode 1).  Wall time with host buffers in and out.
Dense-contraction flops as SURVEY 8(d): (8 N n_rx n_tx + 8 N n_tx N) per (channel, slice)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wofdm_b200 as W
from wofdm_b200 import capi, ofdm_utils as U


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 250
    rng = np.random.default_rng(0)
    chan = (rng.standard_normal((21, C)) + 1j * rng.standard_normal((21, C))) * np.exp(-np.arange(21) / 4)[:, None]
    variants = {"CP": 1, "wtx": 2, "CPwtx": 2, "wrx": 2, "CPwrx": 2, "WOLA": 7, "CPW": 7}
    h = W.Handle([0])
    out = {}
    for mode in (0, 1, 2):
        calls, flops, t0 = 0, 0.0, None
        for rep in range(2):                      # first pass = warm-up (arena growth, module load)
            calls, flops = 0, 0.0
            t0 = time.perf_counter()
            for name, nv in variants.items():
                ttx = 8 if name in U.TX_SYSTEMS else 0
                trx = 10 if name in U.RX_SYSTEMS else 0
                for cp in range(10, 33, 2):
                    s = capi.params_from_name(name, 256, cp, ttx, trx, precision=1)
                    vt0, vr0 = capi.rc_window_tx(s), capi.rc_window_rx(s)
                    M = 1 + -(-(21 - 1 + ttx) // s.stride)
                    for v in range(nv):           # window variants: RC tails scaled by a seeded +-10 %
                        vt = np.clip(vt0 * (1 + 0.02 * v), 0, 1.2)
                        P = h.interf_power(s, vt, vr0, chan, mode=mode)
                        calls += 1
                        flops += (8 * 256 * s.stride * s.n_tx + 8 * 256 * s.n_tx * 256) * C * M
            dt = time.perf_counter() - t0
        out[f"mode{mode}"] = {"calls": calls, "wall_s": dt, "ms_per_call": dt / calls * 1e3,
                              "dense_contraction_TFLOPs": flops / dt / 1e12, "channels": C, "checksum": float(P.sum())}
    print(json.dumps(out))


main()
