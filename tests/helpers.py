"""Shared helpers for the test-suite (oracle side only)."""
import os

import numpy as np

from oracle import wofdm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_ser_golden(name):
    g = np.load(os.path.join(GOLDEN, f"ser_{name}.npz"))
    return {k: g[k] for k in g.files}


def golden_params(g, S, **kw):
    return O.system_params(str(g["name"]), int(g["N"]), int(g["cp"]), int(g["tail_tx"]), int(g["tail_rx"]),
                           S=int(S), **kw)


def golden_windows(g, p):
    """Window pairs in the order the reference evaluates them (opt, then RC)."""
    if str(g["name"]) == "CP":
        return [(np.ones(p.n_tx), np.ones(p.N + p.tail_rx))]
    return [(g["v_tx"], g["v_rx"]), (O.rc_window_tx(p), O.rc_window_rx(p))]


def replay_frames(p, windows, channels, ensemble, snr_arr, seed):
    """Replay np.random in the reference's order and return every frame's
    injected draws: list of dict(snr, chan, sym_idx, noise[w])."""
    pts = O.qam_points(p.bits, 0)
    n = O.noise_len(p, channels.shape[0])
    np.random.seed(seed)
    frames = []
    for snr in snr_arr:
        for c in range(channels.shape[1]):
            for _ in range(ensemble):
                X = np.random.choice(pts, size=(p.N, p.S), replace=True)
                idx = O.hard_decision(X, p.bits, 0)
                noises = []
                for _w in windows:
                    g1 = np.random.randn(n)
                    g2 = np.random.randn(n)
                    noises.append(g1 + 1j * g2)
                frames.append(dict(snr=float(snr), chan=channels[:, c], sym_idx=idx, noise=noises))
    return frames
