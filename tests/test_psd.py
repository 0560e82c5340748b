"""Next row 8f-3: PSD / out-of-band-radiation estimate (python/ofdm_utils/timefreq_simulation.py).

CPU: the oracle restatement against the reference's own estimate_obr outputs (tests/golden/psd.npz, made by
tests/golden/make_golden_psd.py).  GPU: wofdm_psd_estimate with the reference's symbols injected against the same
golden vectors, production draws against the oracle's ensemble average, the timefreq mirror's files and keys."""
import os

import numpy as np
import pytest

import wofdm_b200 as W
from oracle import wofdm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
PTS = np.array([(2 * a - 3) + 1j * (2 * c - 3) for a in range(4) for c in range(4)])     # the reference's `symbols` order


def cases():
    g = np.load(os.path.join(HERE, "golden", "psd.npz"))
    for k in range(int(g["n_cases"])):
        name, cp, ttx, trx = str(g[f"c{k}_name"]), int(g[f"c{k}_cp"]), int(g[f"c{k}_ttx"]), int(g[f"c{k}_trx"])
        p = O.system_params(name, 256, cp, ttx, trx, S=16, bits=4)
        tail = ttx if name in ("wtx", "CPwtx", "WOLA", "CPW") else 0
        sig = {"opt": (g[f"c{k}_win_tx"], tail), "rc": (O.rc_window_tx(p), tail), "cp": (np.ones(p.n_tx), 0)}
        yield k, g, name, p, sig


def test_oracle_reproduces_the_reference_estimate():
    for k, g, name, p, sig in cases():
        X = PTS[g[f"c{k}_idx"].astype(int)]
        for key, (w, tail) in sig.items():
            est = O.psd_estimate(O.psd_tx_stream(256, p.cp, p.cs, tail, w, X, 48), 2048)
            ref = g[f"c{k}_X_{key}"]
            assert np.max(np.abs(est - ref)) < 1e-12 * np.max(ref), (name, key)


def test_analytical_psd_of_the_mirror_matches_the_reference():
    """The closed-form curves S_opt / S_rc / S_cp are host code (no GPU needed): the mirror's analytical_psd against the
    reference's own outputs, and the allocation of estimate_obr (DC and 95 centre bins null, 160 data rows)."""
    from wofdm_b200 import timefreq
    bins = O.psd_data_bins(256, 48)
    assert bins.size == 160 and bins[0] == 1 and bins[79] == 80 and bins[80] == 176 and bins[-1] == 255
    for k, g, name, p, sig in cases():
        ttx, trx = int(g[f"c{k}_ttx"]), int(g[f"c{k}_trx"])
        m = timefreq.wOFDMSystem(name, 256, p.cp, ttx, trx, "/tmp")
        assert m.cs_len == p.cs
        S = m.analytical_psd(200e-9, np.diag(g[f"c{k}_win_tx"]), 48, 2048)
        for got, key in zip(S, ("opt", "rc", "cp")):
            ref = g[f"c{k}_S_{key}"]
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-12 * np.max(ref)), (name, key)


@pytest.fixture(scope="module")
def handle():
    h = W.Handle([0])
    yield h
    h.close()


@pytest.mark.gpu
def test_device_estimate_with_the_reference_symbols(handle):
    """Injected symbols, one record: the device's X_est against the reference's for the optimised, RC and CP signals
    (fp32 chain: 1e-4 of each bin plus 1e-9 of the in-band level, 60 dB below the out-of-band floor)."""
    for k, g, name, p, sig in cases():
        idx = g[f"c{k}_idx"].astype(np.int32).T[None]                    # (1, n_sym, N - 96)
        for key, (w, tail) in sig.items():
            got = handle.psd_estimate(256, p.cp, p.cs, tail, w, sym_idx=idx)
            ref = g[f"c{k}_X_{key}"]
            assert got.shape == (2048,)
            assert np.all(np.abs(got - ref) <= 1e-4 * ref + 1e-9 * np.max(ref)), (name, key, np.max(np.abs(got - ref) / ref))


@pytest.mark.gpu
def test_device_draws_give_the_same_spectrum(handle):
    """Production mode (Philox draws, many records) against the oracle's average over numpy draws: in-band level and
    out-of-band radiation agree within the Monte-Carlo error; the result does not depend on the grid (same seed twice)."""
    p = O.system_params("wtx", 256, 16, 8, 0, S=16, bits=4)
    w = O.perturbed_windows(p, seed=4)[0]
    a = handle.psd_estimate(256, 16, p.cs, 8, w, records=48, seed=5)
    b = handle.psd_estimate(256, 16, p.cs, 8, w, records=48, seed=5)
    assert np.allclose(a, b, rtol=1e-12, atol=0)                         # fp64 atomics: order only
    rng = np.random.default_rng(1)
    ref = np.zeros(2048)
    R = 12
    for _ in range(R):
        X = PTS[rng.integers(0, 16, size=(160, 256))]
        ref += O.psd_estimate(O.psd_tx_stream(256, 16, p.cs, 8, w, X, 48), 2048) / R
    gb = 8 * 48
    inband = slice(gb + 16, 1024 - 16)
    assert abs(a[inband].mean() / ref[inband].mean() - 1) < 0.02
    oob = np.r_[0:gb - 16, 2048 - gb + 16:2048]
    assert abs(a[oob].mean() / ref[oob].mean() - 1) < 0.10
    assert a[oob].mean() < 1e-2 * a[inband].mean()                       # the window does its job
    c = handle.psd_estimate(256, 16, p.cs, 8, w, records=48, seed=6)
    assert not np.array_equal(a, c)


@pytest.mark.gpu
def test_timefreq_mirror(handle, tmp_path):
    """Drop-in timefreq_fun: the reference's tuple and window file in, its three .npz files out (same keys); the
    analytical curves equal the reference's, the estimated OBR of the reference's own draw is a plausible sample."""
    from wofdm_b200 import timefreq
    for k, g, name, p, sig in cases():
        ttx, trx = int(g[f"c{k}_ttx"]), int(g[f"c{k}_trx"])
        wdir = tmp_path / f"win{k}"
        wdir.mkdir()
        if name != "CP":
            x_tx = O.reduce_window_tx(g[f"c{k}_win_tx"], p) if ttx else np.zeros(0)
            x_rx = np.ones(trx // 2 + 1) if trx else np.zeros(0)
            np.save(wdir / f"{name}_{p.cp}.npy", np.concatenate([x_tx, x_rx]))
        # the reference passes tail_tx = 0 for the systems without a Tx window (their stored variable has no Tx part)
        opt, rc, cpd = timefreq.timefreq_fun((name, 256, p.cp, ttx, trx, str(wdir), str(tmp_path)), seed=3, records=4,
                                             handle=handle)
        for d, key in ((opt, "opt"), (rc, "rc"), (cpd, "cp")):
            assert set(d) == {f"X_est_{key}", f"S_{key}", "f_axis", f"obr_{key}", f"mf_band_{key}"}
            assert np.allclose(d[f"S_{key}"], g[f"c{k}_S_{key}"], rtol=1e-9, atol=1e-12 * np.max(g[f"c{k}_S_{key}"])), (name, key)
        got = np.array([opt["obr_opt"], rc["obr_rc"], cpd["obr_cp"]])
        assert np.all(np.abs(got / g[f"c{k}_obr"] - 1) < 0.25), (name, got, g[f"c{k}_obr"])
        z = np.load(tmp_path / "timefreq" / f"opt_{name}_{p.cp}.npz")
        assert np.array_equal(z["X_est_opt"], opt["X_est_opt"]) and z["f_axis"].shape == (2048,)
        assert (tmp_path / "timefreq" / f"rc_{name}_{p.cp}.npz").exists() and (tmp_path / "timefreq" / f"CP_{p.cp}.npz").exists()


@pytest.mark.gpu
def test_psd_rejects_bad_arguments(handle):
    w = np.ones(256 + 16 + 8)
    with pytest.raises(W.WofdmError):
        handle.psd_estimate(512, 16, 8, 8, np.ones(512 + 24))           # N = 256 only
    with pytest.raises(W.WofdmError):
        handle.psd_estimate(256, 16, 8, 8, w, guard_band=128)
    with pytest.raises(W.WofdmError):
        handle.psd_estimate(256, 16, 8, 8, w, n_sym=4)                  # shorter than one slice
    with pytest.raises(W.WofdmError):
        handle.psd_estimate(256, 16, 8, 8, w[:-1])
