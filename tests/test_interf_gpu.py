"""GPU parity tests of K2 (interference power) through the C-ABI.

fp64 tensor path (mode 0) and its Hermitian-form variant (mode 2: impulse responses through the same contraction, then
a quadratic form in the taps per channel): <= 1e-9 relative against (i) the reference's own interf_power outputs on the mean
channel (tests/golden/interf_*.npz), (ii) the dense oracle per channel, (iii) the independent Gram-form
oracle; MATLAB scalar semantics against the restated calculate_interference."""
import numpy as np
import pytest

import wofdm_b200 as W
from oracle import wofdm_oracle as O
from helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    with W.Handle([0]) as h:
        yield h


def to_sys(p):
    return W.SysT(N=p.N, cp=p.cp, cs=p.cs, tail_tx=p.tail_tx, tail_rx=p.tail_rx, rm=p.rm, shift=p.shift,
                  bits=p.bits, S=p.S, noise_norm=0, constellation=0, precision=1)


def rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("name", O.SYSTEMS)
def test_golden_mean_channel(handle, name, mode):
    """The reference evaluates the MEAN impulse response (interf_calc.py:80-83): same call, one column."""
    g = np.load(f"{GOLDEN}/interf_{name}.npz")
    ttx, trx = (0, 0) if name == "CP" else (int(g["tail_tx"]), int(g["tail_rx"]))
    p = O.system_params(name, int(g["N"]), int(g["cp"]), ttx, trx)
    h = g["channels"].mean(axis=1)
    vt, vr = (np.ones(p.n_tx), np.ones(p.N)) if name == "CP" else (g["v_tx"], g["v_rx"])
    P = handle.interf_power(to_sys(p), vt, vr, h[:, None], mode=mode)
    assert P.shape == (1, p.N)
    assert rel(P[0], g["P_opt"]) < 1e-9
    if name != "CP":
        Prc = handle.interf_power(to_sys(p), O.rc_window_tx(p), O.rc_window_rx(p), h[:, None], mode=mode)
        assert rel(Prc[0], g["P_rc"]) < 1e-9


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("name", ["wtx", "WOLA", "CPW", "CPwrx"])
def test_per_channel_batch(handle, name, mode):
    """Every channel realisation at once (the new capability): oracle = the same function per column."""
    g = np.load(f"{GOLDEN}/interf_{name}.npz")
    p = O.system_params(name, 256, 16, int(g["tail_tx"]), int(g["tail_rx"]))
    chans = g["channels"][:, :6]
    P = handle.interf_power(to_sys(p), g["v_tx"], g["v_rx"], chans, mode=mode)
    for c in range(chans.shape[1]):
        want = O.interf_power_dense(p, g["v_tx"], g["v_rx"], chans[:, c])
        assert rel(P[c], want) < 1e-9, c
    want = O.interf_power_gram(p, g["v_tx"], g["v_rx"], chans[:, 0])
    assert rel(P[0], want) < 1e-9


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("N,cp,ttx,trx,L", [(64, 8, 2, 2, 9), (128, 16, 4, 4, 21), (512, 24, 8, 8, 33), (64, 8, 2, 2, 80)])
def test_other_sizes_and_long_channels(handle, N, cp, ttx, trx, L, mode):
    """Other DFT sizes; L = 80 > n_rx gives more than one ISI slice (M = 3)."""
    p = O.system_params("WOLA", N, cp, ttx, trx)
    vt, vr, _, _ = O.perturbed_windows(p, seed=N)
    chans = O.synth_channels(3, L, seed=L)
    assert O.n_slices(p, L) == (3 if L == 80 else 2)
    P = handle.interf_power(to_sys(p), vt, vr, chans, mode=mode)
    for c in range(3):
        assert rel(P[c], O.interf_power_dense(p, vt, vr, chans[:, c])) < 1e-9


@pytest.mark.parametrize("mode", [0, 2])
def test_matlab_scalar_semantics(handle, mode):
    """calculate_interference (main_interference_calculation.m:177-225): ISI slices summed before squaring."""
    p = O.system_params("WOLA", 64, 8, 2, 2)
    vt, vr, _, _ = O.perturbed_windows(p, seed=4)
    chans = O.synth_channels(4, 80, seed=9)          # M = 3: the two conventions differ
    S = handle.interf_power(to_sys(p), vt, vr, chans, mode=mode, scalar=True)
    Pv = handle.interf_power(to_sys(p), vt, vr, chans, mode=mode)
    for c in range(4):
        want = O.interf_power_matlab(p, vt, vr, chans[:, c])
        assert abs(S[c] - want) < 1e-9 * want
        assert abs(Pv[c].sum() - want) > 1e-8 * want      # the python convention (sum of powers) is a different number
    p2 = O.system_params("wtx", 256, 16, 8, 0)
    ch2 = O.synth_channels(2, 21, seed=1)
    S2 = handle.interf_power(to_sys(p2), O.rc_window_tx(p2), O.rc_window_rx(p2), ch2, mode=mode, scalar=True)
    P2 = handle.interf_power(to_sys(p2), O.rc_window_tx(p2), O.rc_window_rx(p2), ch2, mode=mode)
    assert np.allclose(S2, P2.sum(axis=1), rtol=1e-12)     # one ISI slice: identical


@pytest.mark.parametrize("name,N,C,L", [("WOLA", 256, 250, 21), ("CPW", 256, 1000, 21), ("WOLA", 1024, 40, 21), ("wtx", 256, 300, 84)])
def test_hermitian_form_many_channels(handle, name, N, C, L):
    """mode 2 at the sizes it is meant for (BASELINE configs[3]: 250 channels; more channels than taps): the whole batch
    against the direct contraction (mode 0) and a few columns against the dense oracle."""
    sc = N // 256
    ttx, trx = {"WOLA": (8 * sc, 10 * sc), "CPW": (8 * sc, 10 * sc), "wtx": (8 * sc, 0)}[name]
    p = O.system_params(name, N, 16 * sc if L <= 21 else 32 * sc, ttx, trx)
    vt, vr, _, _ = O.perturbed_windows(p, seed=C)
    chans = O.synth_channels(C, L, seed=N + C)
    P0 = handle.interf_power(to_sys(p), vt, vr, chans, mode=0)
    P2 = handle.interf_power(to_sys(p), vt, vr, chans, mode=2)
    assert P2.shape == (C, N)
    assert np.max(np.abs(P2 - P0) / np.max(np.abs(P0), axis=1, keepdims=True)) < 1e-10
    if N == 256:
        for c in (0, C // 2, C - 1):
            assert rel(P2[c], O.interf_power_dense(p, vt, vr, chans[:, c])) < 1e-9, c
    S0 = handle.interf_power(to_sys(p), vt, vr, chans, mode=0, scalar=True)
    S2 = handle.interf_power(to_sys(p), vt, vr, chans, mode=2, scalar=True)
    assert np.allclose(S2, S0, rtol=1e-10)


@pytest.mark.parametrize("L,C", [(1, 3), (2, 1), (5, 70), (21, 1), (88, 2)])
def test_hermitian_form_edge_shapes(handle, L, C):
    """mode 2 with one tap, one channel, fewer channels than taps, the longest channel it takes (88 taps)."""
    p = O.system_params("WOLA", 64, 8, 2, 2)
    vt, vr, _, _ = O.perturbed_windows(p, seed=L)
    chans = O.synth_channels(C, L, seed=C)
    P2 = handle.interf_power(to_sys(p), vt, vr, chans, mode=2)
    for c in range(min(C, 3)):
        want = O.interf_power_dense(p, vt, vr, chans[:, c])
        # (channels inside the prefix leave rounding noise only, ~1e-28 of the signal: absolute floor)
        assert np.max(np.abs(P2[c] - want)) <= 1e-9 * np.max(want) + 1e-24, c
    with pytest.raises(W.WofdmError):
        handle.interf_power(to_sys(p), vt, vr, O.synth_channels(2, 89, seed=1), mode=2)
    with pytest.raises(W.WofdmError):
        handle.interf_power(to_sys(p), vt, vr, chans, mode=3)


@pytest.mark.parametrize("name", ["WOLA", "wtx", "CPW", "CP"])
def test_tf32_split_tensor_path(handle, name):
    """mode 1: tcgen05 kind::tf32 with the 3xTF32 split; fp32-grade accuracy against the fp64 path / oracle.
    Stated bound: |P_tf32 - P_fp64| <= 2e-5 * max(P) + 5e-8 per sub-carrier (split error 2^-21 per product, fp32
    accumulation over K = 544 and 256 columns; the absolute term covers systems whose interference is orders of magnitude
    below the O(1) useful gain the same fp32 accumulators also hold, e.g. CP-OFDM)."""
    g = np.load(f"{GOLDEN}/interf_{name}.npz")
    ttx, trx = (0, 0) if name == "CP" else (int(g["tail_tx"]), int(g["tail_rx"]))
    p = O.system_params(name, 256, 16, ttx, trx)
    vt, vr = (np.ones(p.n_tx), np.ones(p.N)) if name == "CP" else (g["v_tx"], g["v_rx"])
    chans = g["channels"][:, :5]
    P64 = handle.interf_power(to_sys(p), vt, vr, chans, mode=0)
    P32 = handle.interf_power(to_sys(p), vt, vr, chans, mode=1)
    tol = 2e-5 * np.max(P64) + 5e-8
    assert np.max(np.abs(P32 - P64)) < tol
    want = O.interf_power_dense(p, vt, vr, chans[:, 2])
    assert np.max(np.abs(P32[2] - want)) < tol
    S32 = handle.interf_power(to_sys(p), vt, vr, chans, mode=1, scalar=True)
    assert np.allclose(S32, P64.sum(axis=1), rtol=2e-5, atol=256 * 5e-8)


@pytest.mark.parametrize("N,cp,ttx,trx,L", [(512, 32, 16, 20, 21), (1024, 64, 32, 40, 21), (512, 24, 8, 8, 90)])
def test_tf32_split_other_sizes(handle, N, cp, ttx, trx, L):
    """mode 1 for N = 512 / 1024 (several 256-column tiles per slice) and a channel long enough for more ISI rows."""
    p = O.system_params("WOLA", N, cp, ttx, trx)
    vt, vr, _, _ = O.perturbed_windows(p, seed=N)
    chans = O.synth_channels(3, L, seed=N + 1)
    P64 = handle.interf_power(to_sys(p), vt, vr, chans, mode=0)
    P32 = handle.interf_power(to_sys(p), vt, vr, chans, mode=1)
    tol = 2e-5 * np.max(P64) * (N / 256) + 5e-8 * (N / 256)
    assert np.max(np.abs(P32 - P64)) < tol
    with pytest.raises(W.WofdmError):
        p2 = O.system_params("WOLA", 128, 16, 4, 4)
        handle.interf_power(to_sys(p2), O.rc_window_tx(p2), O.rc_window_rx(p2), chans[:9], mode=1)
