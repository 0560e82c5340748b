"""Channel generation: oracle vs the reference's own gen_chan outputs (golden, CPU) and the device
kernel vs the oracle (GPU, through the C-ABI).  fp64 everywhere; tolerance 1e-11 relative (cos of arguments up to
~1e4 rad differs by a few ulp between libraries), written here."""
import os

import numpy as np
import pytest

from oracle import wofdm_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chan_gen.npz"))
CASES = range(int(G["n_cases"]))
REL = 1e-11


def case(k):
    return (str(G[f"c{k}_std"]), int(G[f"c{k}_L"]), int(G[f"c{k}_frames"]), G[f"c{k}_phases"], G[f"c{k}_taps"])


@pytest.mark.parametrize("k", CASES)
def test_oracle_matches_reference_outputs(k):
    std, L, frames, phases, taps = case(k)
    got = O.gen_chan(std, L, float(G["fd"]), float(G["fs"]), float(G["frame"]), frames, phases)
    assert got.shape == taps.shape == (L, frames)
    assert np.abs(got - taps).max() <= 1e-14 * np.abs(taps).max()
    # the reference's normalisation: every path carries ENERGY 10^(dB/10) over the frames of the set
    delays, powers = O.ITUR_CHANNELS[std]
    w = O.gmeds1_waveforms(phases, float(G["fd"]), frames, 1 / float(G["frame"]))
    assert w.shape == (frames, len(delays))


@pytest.mark.gpu
@pytest.mark.parametrize("k", CASES)
def test_device_matches_golden_with_injected_phases(k):
    import wofdm_b200 as W
    std, L, frames, phases, taps = case(k)
    with W.Handle([0]) as h:
        got = h.gen_channels(std, L, float(G["fd"]), float(G["fs"]), float(G["frame"]), no_frames=frames, n_sets=1,
                             phases=phases[None])
        assert got.shape == (L, frames)
        assert np.abs(got - taps).max() <= REL * np.abs(taps).max()
        # several sets in one launch = the same sets one by one
        ph3 = np.stack([phases, phases[::-1], 0.5 * phases])
        many = h.gen_channels(std, L, float(G["fd"]), float(G["fs"]), float(G["frame"]), no_frames=frames, n_sets=3, phases=ph3)
        for i in range(3):
            ref = O.gen_chan(std, L, float(G["fd"]), float(G["fs"]), float(G["frame"]), frames, ph3[i])
            assert np.abs(many[:, i * frames:(i + 1) * frames] - ref).max() <= REL * np.abs(ref).max()


@pytest.mark.gpu
def test_long_record_is_spread_over_blocks():
    """no_frames beyond one CTA's share (two-pass energy): still the oracle's values, and the record's energy per path
    is the profile power."""
    import wofdm_b200 as W
    rng = np.random.default_rng(3)
    phases = rng.standard_normal((1, 6, 21, 2))
    frames = 1700
    with W.Handle([0]) as h:
        got = h.gen_channels("vehicularA", 21, float(G["fd"]), float(G["fs"]), float(G["frame"]), no_frames=frames, n_sets=1,
                             phases=phases)
    ref = O.gen_chan("vehicularA", 21, float(G["fd"]), float(G["fs"]), float(G["frame"]), frames, phases[0])
    assert got.shape == ref.shape == (21, frames)
    assert np.abs(got - ref).max() <= REL * np.abs(ref).max()


@pytest.mark.gpu
def test_production_draws_and_driver_mirror(tmp_path):
    """On-device Philox phases: deterministic per seed, independent sets, and -- one frame per set, as the reference's
    driver stores channels -- every path has exactly its profile power with a uniform phase, so the mean tap energy
    equals sum_p P_p * sum_l sinc^2."""
    import wofdm_b200 as W
    from wofdm_b200 import channel_model as CM
    with W.Handle([0]) as h:
        a = CM.gen_channel_set("vehicularA", 4000, str(tmp_path), seed=5, handle=h)
        b = CM.gen_channel_set("vehicularA", 4000, None, seed=5, handle=h)
        c = CM.gen_channel_set("vehicularA", 4000, None, seed=6, handle=h)
        assert a.shape == (21, 4000) and a.dtype == np.complex128
        assert np.array_equal(a, b) and not np.array_equal(a, c)
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "vehicularA.npy")), a)
        delays, powers = O.ITUR_CHANNELS["vehicularA"]
        axis = np.linspace(-10, 11, 21)
        sinc = np.sinc(np.asarray(delays)[None, :] / 200e-9 - axis[:, None])
        want = (sinc ** 2 * (10 ** (np.asarray(powers) / 10))[None, :]).sum()
        got = (np.abs(a) ** 2).sum(axis=0).mean()
        assert abs(got - want) < 0.05 * want, (got, want)
        # distinct sets are uncorrelated
        r = np.abs(np.vdot(a[:, :2000].ravel(), a[:, 2000:].ravel())) / np.linalg.norm(a[:, :2000]) / np.linalg.norm(a[:, 2000:])
        assert r < 0.05
        one = CM.gen_chan("vehicularB", 33, 185.0, 5e6, 8.192e-4, 50, seed=1, handle=h)
        assert one.shape == (33, 50)
        with pytest.raises(W.WofdmError):
            h.gen_channels("vehicularZ", 21, 185.0, 5e6, 8.192e-4)
