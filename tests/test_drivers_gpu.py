"""GPU tests of the reference-facing host mirror (w-ofdm-optimization_b200/ofdm_utils.py) and of production mode
(on-device RNG): SER/BER must fall inside the 95 % confidence interval of the oracle at every SNR point
(north star), file inputs/outputs keep the reference's schema."""
import os

import numpy as np
import pytest

import wofdm_b200 as W
from wofdm_b200 import ofdm_utils as U
from oracle import wofdm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    with W.Handle([0]) as h:
        U.set_handle(h)
        yield h
        U.set_handle(None)


def oracle_frame_sers(p, vt, vr, chans, snr_arr, frames_per_chan, seed, bits_mode=False):
    """per-frame error ratios from the numpy oracle with its own seeded draws -> mean and 95 % CI"""
    rng = np.random.default_rng(seed)
    L = chans.shape[0]
    n = O.noise_len(p, L)
    out = []
    for snr in snr_arr:
        vals = []
        for c in range(chans.shape[1]):
            for _ in range(frames_per_chan):
                idx = rng.integers(0, 1 << p.bits, size=(p.N, p.S))
                nz = rng.standard_normal(n) + 1j * rng.standard_normal(n)
                r = O.frame_chain_structured(p, vt, vr, chans[:, c], float(snr), idx, nz)
                vals.append(r.bit_err / (p.N * p.bits * (p.S - 1)) if bits_mode else r.sym_err / (p.N * (p.S - 1)))
        v = np.array(vals)
        out.append((v.mean(), 1.96 * v.std(ddof=1) / np.sqrt(len(v))))
    return out


@pytest.mark.parametrize("name,conv,nn", [("wtx", 0, 0), ("WOLA", 1, 1), ("CP", 0, 0)])
def test_production_ser_within_oracle_ci(handle, name, conv, nn):
    ttx = 8 if name in U.TX_SYSTEMS else 0
    trx = 10 if name in U.RX_SYSTEMS else 0
    p = O.system_params(name, 256, 16, ttx, trx, S=16, constellation=conv, noise_norm=nn)
    vt, vr, _, _ = O.perturbed_windows(p, seed=8)
    chans = O.synth_channels(4, 21, seed=12)
    snr = np.array([0.0, 10.0, 20.0, 30.0])
    ref = oracle_frame_sers(p, vt, vr, chans, snr, 40, seed=77, bits_mode=(conv == 1))
    s = W.SysT(N=p.N, cp=p.cp, cs=p.cs, tail_tx=p.tail_tx, tail_rx=p.tail_rx, rm=p.rm, shift=p.shift, bits=p.bits,
               S=p.S, noise_norm=nn, constellation=conv, precision=0)
    r = handle.ber_run(s, vt, vr, chans, snr, 2000, seed=5)
    got = r["bit_err"] / r["bit_tot"] if conv == 1 else r["sym_err"] / r["sym_tot"]
    # 95 % CI of the oracle's own Monte-Carlo estimate (160 frames per point).  Twelve two-sided 95 % checks
    # would fail by chance every other run, so: every point inside the 99.9 % interval, >= 3 of 4 inside 95 %.
    inside = 0
    for k, (mean, ci) in enumerate(ref):
        assert abs(got[k] - mean) <= ci * (3.29 / 1.96) + 1e-4, (name, snr[k], got[k], mean, ci)
        inside += abs(got[k] - mean) <= ci + 1e-4
    assert inside >= 3, (name, got, ref)


def test_simulation_fun_files(handle, tmp_path):
    """Same 11-tuple, same input files (<sys>_<cp>.npy reduced tails, channels .npy), same output files."""
    win_dir, out_dir = tmp_path / "optimized_windows", tmp_path / "simulation_results"
    os.makedirs(win_dir)
    chans = O.synth_channels(3, 21, seed=2)
    np.save(tmp_path / "vehicularA.npy", chans)
    snr = np.arange(-21, 51, 12)
    for name in ("WOLA", "wtx", "wrx", "CP"):
        ttx = 8 if name in U.TX_SYSTEMS else 0
        trx = 10 if name in U.RX_SYSTEMS else 0
        p = O.system_params(name, 256, 16, ttx, trx)
        _, _, xt, xr = O.perturbed_windows(p, seed=1)
        if name in ("WOLA", "CPW"):
            np.save(win_dir / f"{name}_16.npy", np.concatenate([xt, xr]))
        elif name == "wtx":
            np.save(win_dir / f"{name}_16.npy", xt)
        elif name == "wrx":
            np.save(win_dir / f"{name}_16.npy", xr)
        U.simulation_fun((name, 256, 16, ttx, trx, str(tmp_path / "vehicularA.npy"), str(win_dir), 40, snr, 16,
                          str(out_dir)))
        files = [f"CP_16.npy"] if name == "CP" else [f"opt_{name}_16.npy", f"rc_{name}_16.npy"]
        for f in files:
            ser = np.load(out_dir / "ser" / f)
            assert ser.shape == (len(snr),) and ser.dtype == np.float64
            assert ser[0] > 0.8 and np.all(np.diff(ser) <= 0.02) and ser[-1] < 0.2


def test_matlab_signatures(handle):
    """run_simulation / calculate_interference with the MATLAB argument lists (diagonal-matrix windows)."""
    p = O.system_params("WOLA", 256, 16, 8, 10, constellation=1, noise_norm=1)
    vt, vr = O.rc_window_tx(p), O.rc_window_rx(p)
    chans = O.synth_channels(5, 21, seed=3)
    bers = [U.run_simulation(50, 16, 4, 256, 16, p.cs, np.diag(vt), chans[:, 0], snr, 8, 10, np.diag(vr), p.rm, p.shift)
            for snr in (0.0, 15.0, 30.0)]
    assert 0.2 < bers[0] < 0.5 and bers[0] > bers[1] > bers[2]
    got = U.calculate_interference(16, "WOLA", np.diag(vt), np.diag(vr), 256, 8, 10, chans.T)
    want = O.interf_power_matlab(p, vt, vr, chans.mean(axis=1))
    assert abs(got - want) < 1e-9 * want


def test_interf_power_mirror(handle, tmp_path):
    chans = O.synth_channels(6, 21, seed=4)
    path = tmp_path / "vehicularA.npy"
    np.save(path, chans)
    p = O.system_params("CPW", 256, 16, 8, 10)
    vt, vr, _, _ = O.perturbed_windows(p, seed=6)
    Po, Prc = U.interf_power("CPW", [np.diag(vt), np.diag(vr)], 256, 16, 8, 10, channel_path=str(path))
    assert np.allclose(Po, O.interf_power_dense(p, vt, vr, chans.mean(axis=1)), rtol=1e-9)
    assert np.allclose(Prc, O.interf_power_dense(p, O.rc_window_tx(p), O.rc_window_rx(p), chans.mean(axis=1)), rtol=1e-9)
    Pc = U.interf_power("CP", [None, None], 256, 16, 0, 0, channel_path=str(path))
    pc = O.system_params("CP", 256, 16, 0, 0)
    assert np.allclose(Pc, O.interf_power_dense(pc, np.ones(pc.n_tx), np.ones(256), chans.mean(axis=1)), rtol=1e-9)
    Pall, _ = U.interf_power("CPW", [vt, vr], 256, 16, 8, 10, channel_path=str(path), per_channel=True)
    assert Pall.shape == (6, 256)


@pytest.mark.parametrize("name", O.SYSTEMS)
def test_production_ser_against_the_reference_itself(handle, name):
    """North star: production-mode SER inside the reference's confidence interval at every SNR point.  The golden
    holds SER curves from the REFERENCE's own jitted Monte-Carlo loops (tests/golden/make_golden_stat.py: 4 channels x
    ensemble 150 = 600 frames per point).  The device repeats that experiment 24 times with different seeds (its own
    Philox draws); the reference's estimate must be a plausible draw of that sampling distribution (|z| <= 3.5 at all
    5 x 2 points) and agree with the device's long-run mean within 3.5 standard errors."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ser_statistical.npz"))
    ttx = 8 if name in U.TX_SYSTEMS else 0
    trx = 10 if name in U.RX_SYSTEMS else 0
    p = O.system_params(name, int(g["N"]), int(g["cp"]), ttx, trx, S=int(g["S"]))
    s = W.SysT(N=p.N, cp=p.cp, cs=p.cs, tail_tx=p.tail_tx, tail_rx=p.tail_rx, rm=p.rm, shift=p.shift, bits=4, S=p.S,
               noise_norm=0, constellation=0, precision=0)
    windows = [(g[f"vtx_{name}"], g[f"vrx_{name}"])]
    if name != "CP":
        windows.append((O.rc_window_tx(p), O.rc_window_rx(p)))
    ref = g[f"ser_{name}"]
    for w, (vt, vr) in enumerate(windows):
        runs = np.array([(lambda r: r["sym_err"] / r["sym_tot"])(
            handle.ber_run(s, vt, vr, g["channels"], g["snr"], int(g["ensemble"]), seed=1000 + k, variant=w)) for k in range(24)])
        mean, std = runs.mean(axis=0), runs.std(axis=0, ddof=1)
        z = (ref[w] - mean) / np.sqrt(std ** 2 * (1 + 1 / 24))
        assert np.all(np.abs(z) <= 3.5), (name, w, z, ref[w], mean)


@pytest.mark.parametrize("name,cp,bits,S", [("wtx", 16, 4, 16), ("WOLA", 22, 4, 16), ("CP", 16, 2, 16), ("CPW", 16, 6, 16), ("WOLA", 16, 4, 3)])
def test_channel_mask_variant_replay(handle, name, cp, bits, S):
    """matlab/main_channel_mask.m: guard band N/4 + DFT-domain RC mask.  The production counters of the
    masked chain, replayed through the oracle's line-by-line restatement (frame_chain_masked) from the exported draws;
    the unmasked counters of the same symbols through frame_chain_structured.  (MATLAB-only path: parity unpinned.)"""
    ttx = 8 if name in U.TX_SYSTEMS else 0
    trx = 10 if name in U.RX_SYSTEMS else 0
    p = O.system_params(name, 256, cp, ttx, trx, S=S, bits=bits, noise_norm=1, constellation=1, guard=64)
    vt, vr, _, _ = O.perturbed_windows(p, seed=cp)
    s = W.SysT(N=p.N, cp=p.cp, cs=p.cs, tail_tx=p.tail_tx, tail_rx=p.tail_rx, rm=p.rm, shift=p.shift, bits=bits, S=S,
               noise_norm=1, constellation=1, precision=0, guard=64)
    chans = O.synth_channels(2, 21, seed=cp)
    snr = np.array([14.0, 32.0])
    ens = 2
    plain = handle.ber_run(s, vt, vr, chans, snr, ens, seed=5, variant=0)
    masked = handle.ber_run_masked(s, vt, vr, chans, snr, ens, seed=5, variant=1, roll_off=10)
    F = len(snr) * 2 * ens
    sym, nz0 = handle.ber_draws(s, 21, 5, 0, np.arange(F))
    _, nz1 = handle.ber_draws(s, 21, 5, 1, np.arange(F))
    want = np.zeros((2, len(snr), 2), dtype=np.int64)
    for f in range(F):
        c, si = (f // ens) % 2, f // (ens * 2)
        a = O.frame_chain_structured(p, vt, vr, chans[:, c], snr[si], sym[f].T, nz0[f])
        b = O.frame_chain_masked(p, vt, vr, chans[:, c], snr[si], sym[f].T, nz1[f], roll_off=10)
        want[0, si] += (a.sym_err, a.bit_err)
        want[1, si] += (b.sym_err, b.bit_err)
    assert np.all(masked["sym_tot"] == 2 * ens * 128 * (S - 1)) and np.all(masked["bit_tot"] == masked["sym_tot"] * bits)
    for res, w in ((plain, want[0]), (masked, want[1])):
        assert np.all(np.abs(res["sym_err"] - w[:, 0]) <= 3 + 0.003 * w[:, 0]), (res["sym_err"], w[:, 0])
        assert np.all(np.abs(res["bit_err"] - w[:, 1]) <= 5 + 0.003 * w[:, 1]), (res["bit_err"], w[:, 1])
    # the Tx side is the dense tensor-core product (mask_gemm.cu); the per-symbol FFT kernel (mask_kernel.cuh) is the same
    # linear map evaluated another way: counters agree up to decisions on a boundary
    os.environ["WOFDM_MASK_FFT"] = "1"
    try:
        fft = handle.ber_run_masked(s, vt, vr, chans, snr, ens, seed=5, variant=1, roll_off=10)
    finally:
        del os.environ["WOFDM_MASK_FFT"]
    assert np.all(np.abs(fft["sym_err"] - masked["sym_err"]) <= 2), (fft["sym_err"], masked["sym_err"])
    assert np.all(np.abs(fft["bit_err"] - masked["bit_err"]) <= 3), (fft["bit_err"], masked["bit_err"])
    # the MATLAB-signature mirror returns [berMasked, ber]
    bm, b0 = U.run_sim_mc(3, p.cp, p.cs, ttx, trx, np.diag(vt), np.diag(vr), chans[:, 0], 20.0, 64, p.rm, p.shift, 256, bits,
                          16, 10, seed=1, handle=handle)
    assert 0.0 <= b0 < 0.5 and 0.0 <= bm < 0.5


@pytest.mark.parametrize("name,N,cp,ttx,trx,bits,guard,S", [("CP", 128, 8, 0, 0, 2, 0, 16), ("CPW", 512, 32, 16, 20, 4, 128, 16),
                                                             ("wtx", 256, 16, 8, 0, 6, 0, 16), ("WOLA", 256, 30, 8, 10, 4, 64, 7),
                                                             ("CPwtx", 256, 10, 8, 0, 4, 100, 2)])
def test_channel_mask_product_against_fft_kernel(handle, tmp_path, name, N, cp, ttx, trx, bits, guard, S):
    """The masked Tx stream two ways: the dense tensor-core product (mask_gemm.cu: the matrices of main_channel_mask.m:384-417
    multiplied out once, fp16 hi + lo against the exact lattice points) and the per-symbol FFT kernel (mask_kernel.cuh).
    Streams of the first frames agree to fp32 rounding; the K1 kernel that gathers from the product's output itself gives
    the counters of the assembled stream bit for bit; counters of the two Tx kernels differ by boundary decisions only."""
    s = W.params_from_name(name, N, cp, ttx, trx, bits=bits, S=S, noise_norm=1, constellation=1, guard=guard)
    vt, vr = W.capi.rc_window_tx(s), W.capi.rc_window_rx(s)
    chans = O.synth_channels(6, 21, seed=N + cp)
    snr = np.array([6.0, 18.0, 30.0, 42.0])
    ens = 3
    res, dumps = {}, {}
    try:
        for mode in ("fft", "gemm"):
            os.environ["WOFDM_MASK_FFT"] = "1" if mode == "fft" else "0"
            os.environ["WOFDM_MASK_DUMP"] = str(tmp_path / (mode + ".bin"))
            res[mode] = handle.ber_run_masked(s, vt, vr, chans, snr, ens, seed=9, variant=1, roll_off=10)
            dumps[mode] = np.fromfile(tmp_path / (mode + ".bin"), dtype=np.float32)
            del os.environ["WOFDM_MASK_DUMP"]
        res["gather"] = handle.ber_run_masked(s, vt, vr, chans, snr, ens, seed=9, variant=1, roll_off=10)
    finally:
        os.environ.pop("WOFDM_MASK_FFT", None)
        os.environ.pop("WOFDM_MASK_DUMP", None)
    a, b = dumps["fft"], dumps["gemm"]
    assert a.size == b.size == 2 * 4 * (ttx + S * (N + cp + s.cs - ttx))
    assert np.abs(a - b).max() <= 8e-6 * np.abs(a).max(), (np.abs(a - b).max(), np.abs(a).max())    # (observed: <= 2.8e-6; both sides are fp32)
    for k in ("bit_err", "sym_err", "bit_tot", "sym_tot"):
        assert np.array_equal(res["gather"][k], res["gemm"][k]), k
    assert np.all(np.abs(res["fft"]["sym_err"] - res["gemm"]["sym_err"]) <= 2), (res["fft"]["sym_err"], res["gemm"]["sym_err"])
    assert np.all(np.abs(res["fft"]["bit_err"] - res["gemm"]["bit_err"]) <= 4), (res["fft"]["bit_err"], res["gemm"]["bit_err"])
    assert res["gemm"]["sym_err"][0] > res["gemm"]["sym_err"][-1]


def test_channel_mask_product_over_several_batches(handle):
    """More frames than one batch of the mask product holds (16 384): the last, shorter batch reuses the operand buffers."""
    s = W.params_from_name("wtx", 256, 16, 8, 0, bits=4, S=16, noise_norm=1, constellation=1, guard=64)
    vt, vr = W.capi.rc_window_tx(s), W.capi.rc_window_rx(s)
    chans = O.synth_channels(25, 21, seed=3)
    snr = np.linspace(0.0, 40.0, 30)
    ens = 24                                                 # 18 000 frames
    got = handle.ber_run_masked(s, vt, vr, chans, snr, ens, seed=11, variant=1, roll_off=10)
    os.environ["WOFDM_MASK_FFT"] = "1"
    try:
        ref = handle.ber_run_masked(s, vt, vr, chans, snr, ens, seed=11, variant=1, roll_off=10)
    finally:
        del os.environ["WOFDM_MASK_FFT"]
    assert np.all(got["sym_tot"] == 25 * ens * 128 * 15)
    assert np.all(np.abs(got["sym_err"] - ref["sym_err"]) <= 3 + 2e-5 * ref["sym_err"]), (got["sym_err"], ref["sym_err"])
    assert np.all(np.abs(got["bit_err"] - ref["bit_err"]) <= 5 + 2e-5 * ref["bit_err"]), (got["bit_err"], ref["bit_err"])
    again = handle.ber_run_masked(s, vt, vr, chans, snr, ens, seed=11, variant=1, roll_off=10)
    assert np.array_equal(again["bit_err"], got["bit_err"]) and np.array_equal(again["sym_err"], got["sym_err"])


def test_channel_mask_long_frames(handle):
    """Frames of 48 symbols with every sub-carrier active: the mask product's symbol kernel needs more than the default 48 KB
    of shared memory, K1 is the staged kernel (more symbols than one Tx pass).  Replayed through the oracle like the
    reference-sized frames of test_channel_mask_variant_replay."""
    S, bits = 48, 4
    p = O.system_params("wtx", 256, 16, 8, 0, S=S, bits=bits, noise_norm=1, constellation=1, guard=0)
    vt, vr, _, _ = O.perturbed_windows(p, seed=1)
    s = W.SysT(N=p.N, cp=p.cp, cs=p.cs, tail_tx=p.tail_tx, tail_rx=p.tail_rx, rm=p.rm, shift=p.shift, bits=bits, S=S,
               noise_norm=1, constellation=1, precision=0, guard=0)
    chans = O.synth_channels(2, 21, seed=6)
    snr = np.array([12.0, 28.0])
    masked = handle.ber_run_masked(s, vt, vr, chans, snr, 1, seed=5, variant=1, roll_off=10)
    F = len(snr) * 2
    sym, nz1 = handle.ber_draws(s, 21, 5, 1, np.arange(F))
    want = np.zeros((len(snr), 2), dtype=np.int64)
    for f in range(F):
        c, si = f % 2, f // 2
        b = O.frame_chain_masked(p, vt, vr, chans[:, c], snr[si], sym[f].T, nz1[f], roll_off=10)
        want[si] += (b.sym_err, b.bit_err)
    assert np.all(masked["sym_tot"] == 2 * 256 * (S - 1))
    assert np.all(np.abs(masked["sym_err"] - want[:, 0]) <= 3 + 0.003 * want[:, 0]), (masked["sym_err"], want[:, 0])
    assert np.all(np.abs(masked["bit_err"] - want[:, 1]) <= 5 + 0.003 * want[:, 1]), (masked["bit_err"], want[:, 1])
