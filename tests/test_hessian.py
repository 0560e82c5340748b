"""Window-optimisation Hessian: the Gram-form oracle against the REFERENCE's own gen_hessian outputs
(golden, N = 64, CPU) and the device path against the oracle through the C-ABI (GPU).  fp64; 1e-9 relative to the
largest entry (the reference's loop-built DFT matrices carry ~1e-12 themselves)."""
import os
import time

import numpy as np
import pytest

from oracle import wofdm_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hessian.npz"))
SYSTEMS = ("wtx", "CPwtx", "wrx", "CPwrx", "WOLA", "CPW")


def tails(name, ttx, trx):
    return (ttx if name in ("wtx", "CPwtx", "WOLA", "CPW") else 0, trx if name in ("wrx", "CPwrx", "WOLA", "CPW") else 0)


@pytest.mark.parametrize("name", SYSTEMS)
def test_oracle_matches_reference_hessian(name):
    a, b = tails(name, int(G["tail_tx"]), int(G["tail_rx"]))
    p = O.system_params(name, int(G["N"]), int(G["cp"]), a, b)
    H = O.window_hessian(p, G["h"])
    ref = G[f"H_{name}"]
    assert H.shape == ref.shape == ((b // 2 + 1) * (a + 1),) * 2
    assert np.abs(H - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > -1e-9 * np.abs(H).max()     # a Gram matrix
    # the quadratic form IS the interference power of the expanded window (summed over sub-carriers, the ISI slices
    # summed before squaring: interf_power_matlab)
    rng = np.random.default_rng(1)
    xt = np.concatenate([[1.0], rng.uniform(0, 1, a)])
    xr = np.concatenate([[1.0], rng.uniform(0.5, 1, b // 2)])
    x = np.outer(xr, xt).ravel()
    want = O.interf_power_matlab(p, O.expand_window_tx(xt, p), O.expand_window_rx(xr, p), G["h"])
    assert abs(0.5 * x @ H @ x - want) <= 1e-9 * want


@pytest.mark.gpu
@pytest.mark.parametrize("name", SYSTEMS)
def test_device_matches_golden_and_oracle(name):
    import wofdm_b200 as W
    from wofdm_b200 import optimizers as OPT
    a, b = tails(name, int(G["tail_tx"]), int(G["tail_rx"]))
    with W.Handle([0]) as h:
        if name in ("wtx", "CPwtx"):
            opt = OPT.OptimizerTx(name, int(G["N"]), int(G["cp"]), a, handle=h)
        elif name in ("wrx", "CPwrx"):
            opt = OPT.OptimizerRx(name, int(G["N"]), int(G["cp"]), b, handle=h)
        else:
            opt = OPT.OptimizerTxRx(name, int(G["N"]), int(G["cp"]), a, b, handle=h)
        H = opt.gen_hessian(opt.calculate_chann_matrices(G["h"]))
        ref = G[f"H_{name}"]
        assert np.abs(H - ref).max() <= 1e-9 * np.abs(ref).max()
        # settingsData sizes (N = 256, tails 8 / 10, 21 taps): against the oracle's closed form
        a2, b2 = tails(name, 8, 10)
        p = O.system_params(name, 256, 16, a2, b2)
        hch = O.synth_channels(8, 21, seed=4).mean(axis=1)
        s = W.params_from_name(name, 256, 16, a2, b2)
        h.window_hessian(s, hch)
        t0 = time.perf_counter()
        Hd = h.window_hessian(s, hch)
        dt = time.perf_counter() - t0
        Ho = O.window_hessian(p, hch)
        assert np.abs(Hd - Ho).max() <= 1e-9 * np.abs(Ho).max()
        print(f"{name}: {Hd.shape[0]} variables, device call {dt * 1e3:.2f} ms")


@pytest.mark.parametrize("name,side", [("WOLA", "tx"), ("WOLA", "rx"), ("wtx", "tx"), ("CPwrx", "rx")])
def test_matlab_quad_objective_oracle(name, side):
    """quad_objective_tx / _rx (matlab/window_optimization.m:596-680; MATLAB only: parity unpinned): the three loops as
    written against their vector form, and against the identity the device path uses -- with the full window as the
    variable, H = alpha H_ici + (1 - alpha) diag(row sums of H_isi), H_ici / H_isi the Gram parts of window_hessian."""
    a, b = tails(name, 2, 2)
    p = O.system_params(name, 16, 4, a, b)
    vt, vr, _, _ = O.perturbed_windows(p, seed=3)
    hch = O.synth_channels(1, 5, seed=2)[:, 0]
    fixed = vr if side == "tx" else vt
    Hl = O.quad_objective_matlab(p, side, fixed, hch, 0.3, loops=True)
    Hv = O.quad_objective_matlab(p, side, fixed, hch, 0.3)
    n = p.n_tx if side == "tx" else p.N + p.tail_rx
    assert Hl.shape == (n, n) and np.allclose(Hl, Hl.T)
    assert np.abs(Hl - Hv).max() <= 1e-12 * np.abs(Hl).max()
    x0, xs = [], []
    for i in range(n):
        e = np.zeros(n)
        e[i] = 1.0
        A = O.interf_matrices(p, e, fixed, hch) if side == "tx" else O.interf_matrices(p, fixed, e, hch)
        x0.append((A[0] - np.diag(np.diag(A[0]))).ravel())
        xs.append(A[1:].sum(axis=0).ravel())
    x0, xs = np.array(x0), np.array(xs)
    Hc, Hs = 2.0 * (x0 @ x0.conj().T).real, 2.0 * (xs @ xs.conj().T).real
    assert np.abs(Hl - (0.3 * Hc + 0.7 * np.diag(Hs.sum(axis=1)))).max() <= 1e-12 * np.abs(Hl).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,cp,ttx,trx,L", [("WOLA", 64, 8, 2, 2, 9), ("wtx", 64, 8, 4, 0, 9), ("CPW", 256, 16, 8, 10, 21)])
def test_device_matlab_quad_objective(name, N, cp, ttx, trx, L):
    """The MATLAB-signature mirrors (full window as the variable) through wofdm_window_hessian_parts against the oracle."""
    import wofdm_b200 as W
    from wofdm_b200 import ofdm_utils as U
    p = O.system_params(name, N, cp, ttx, trx)
    vt, vr, _, _ = O.perturbed_windows(p, seed=N)
    hch = O.synth_channels(4, L, seed=5).mean(axis=1)
    with W.Handle([0]) as h:
        t0 = time.perf_counter()
        HTx = U.quad_objective_tx(np.diag(vr), N, trx, cp, name, ttx, hch, 0.5, handle=h)
        dt = time.perf_counter() - t0
        want = O.quad_objective_matlab(p, "tx", vr, hch, 0.5)
        assert HTx.shape == (p.n_tx, p.n_tx)
        assert np.abs(HTx - want).max() <= 1e-9 * np.abs(want).max()
        print(f"{name} N={N}: HTx {HTx.shape[0]} variables, device call {dt * 1e3:.1f} ms")
        if N == 64:
            HRx = U.quad_objective_rx(vt, N, trx, cp, name, ttx, hch, 0.25, handle=h)
            want = O.quad_objective_matlab(p, "rx", vt, hch, 0.25)
            assert np.abs(HRx - want).max() <= 1e-9 * np.abs(want).max()
            Hc, Hs = h.window_hessian_parts(W.params_from_name(name, N, cp, ttx, trx), hch,
                                            *[np.array(x) for x in O.window_basis(p)])
            assert np.abs(Hc + Hs - O.window_hessian(p, hch)).max() <= 1e-9 * np.abs(Hc + Hs).max()   # parts on the reduced bases = wofdm_window_hessian


@pytest.mark.gpu
def test_bad_system_names_raise():
    from wofdm_b200 import optimizers as OPT
    with pytest.raises(ValueError):
        OPT.OptimizerTx("WOLA", 256, 16, 8)
