"""Runs only in the builder container (needs /root/reference): the oracle against the reference executed live,
on draws the committed goldens do not contain.  The GPU box skips this file."""
import os
import sys

import numpy as np
import pytest

from oracle import wofdm_oracle as O

REF = "/root/reference/python"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        from ofdm_utils.wofdm_simulation import wOFDMSystem
        from ofdm_utils.transmitter import gen_rc_window_tx
        from ofdm_utils.receiver import gen_rc_window_rx
        yield wOFDMSystem, gen_rc_window_tx, gen_rc_window_rx
    finally:
        sys.path.remove(REF)


@pytest.mark.parametrize("name,seed", [("WOLA", 101), ("CPwtx", 102), ("wrx", 103)])
def test_live_replay_is_bit_identical(ref, name, seed):
    wOFDMSystem, rc_tx, rc_rx = ref
    N, cp = 256, 22
    ttx = 8 if name in ("CPW", "WOLA", "CPwtx", "wtx") else 0
    trx = 10 if name in ("CPW", "WOLA", "CPwrx", "wrx") else 0
    p = O.system_params(name, N, cp, ttx, trx, S=3)
    vt, vr, _, _ = O.perturbed_windows(p, seed=seed)
    s = wOFDMSystem(name, N, cp, ttx, trx, "/tmp/unused")
    tx = np.diag(vt) @ s.add_red_mat @ s.idft_mat
    tx_rc = rc_tx(N, cp, s.cs_len, ttx) @ s.add_red_mat @ s.idft_mat
    rx = s.dft_mat @ s.circ_shift_mat @ s.overlap_add_mat @ np.diag(vr) @ s.rm_red_mat
    rx_rc = s.dft_mat @ s.circ_shift_mat @ s.overlap_add_mat @ rc_rx(N, trx) @ s.rm_red_mat
    chans = O.synth_channels(2, 21, seed=seed)
    snr = np.array([3.0, 21.0])
    np.random.seed(seed)
    a, b = wOFDMSystem._wOFDMSystem__run_sim_mc.py_func(tx, tx_rc, rx, rx_rc, 3, chans, 2, snr, ttx, True)
    np.random.seed(seed)
    ser = O.ser_sweep_replay(p, [(vt, vr), (O.rc_window_tx(p), O.rc_window_rx(p))], chans, 2, snr, dense=True)
    assert np.array_equal(ser[0], a) and np.array_equal(ser[1], b)
