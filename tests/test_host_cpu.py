"""CPU tests (no GPU needed): the C-ABI library loads and exports every symbol include/wofdm.h declares, the host
logic (parameter table, windows, validation, sharding) agrees with the oracle, and the MEX gateway compiles."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import wofdm_b200 as W
from wofdm_b200 import capi, sharding
from oracle import wofdm_oracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(REPO, "include", "wofdm.h")).read()
    return sorted(set(re.findall(r"WOFDM_API\s+[\w\s\*]+?\b(wofdm_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = header_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), n
        assert n in capi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(capi.SIGNATURES) == names
    assert lib.wofdm_version() == 100


def test_no_cpu_fallback():
    n = ctypes.c_int(-1)
    rc = capi.load().wofdm_device_count(ctypes.byref(n))
    if rc == capi.OK and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(W.WofdmError) as e:
        W.Handle()
    assert e.value.code == capi.ENODEV


@pytest.mark.parametrize("name", O.SYSTEMS)
def test_parameter_table_and_windows(name):
    for cp in (10, 16, 32):
        ttx = 8 if name in ("CPW", "WOLA", "CPwtx", "wtx") else 0
        trx = 10 if name in ("CPW", "WOLA", "CPwrx", "wrx") else 0
        p = O.system_params(name, 256, cp, ttx, trx)
        s = capi.params_from_name(name, 256, cp, ttx, trx)
        assert (s.cs, s.rm, s.shift, s.n_tx, s.stride) == (p.cs, p.rm, p.shift, p.n_tx, p.stride)
        assert np.allclose(capi.rc_window_tx(s), O.rc_window_tx(p), atol=1e-15)
        assert np.allclose(capi.rc_window_rx(s), O.rc_window_rx(p), atol=1e-15)
        vt, vr, xt, xr = O.perturbed_windows(p, seed=cp)
        assert np.array_equal(capi.expand_window_tx(s, xt), vt)
        assert np.array_equal(capi.expand_window_rx(s, xr), vr)
    with pytest.raises(W.WofdmError):
        capi.params_from_name("OFDM?", 256, 16, 8, 10)
    with pytest.raises(W.WofdmError):
        capi.expand_window_tx(s, np.ones(3))


def test_sharding_partition():
    n_snr, C, ens = 5, 7, 3
    total = n_snr * C * ens
    for world in (1, 2, 3, 8):
        ids = np.concatenate([sharding.frame_ids(n_snr, C, ens, (r, world)) for r in range(world)])
        assert sorted(ids.tolist()) == list(range(total))
        per = sum(sharding.frames_per_snr(n_snr, C, ens, (r, world)) for r in range(world))
        assert np.array_equal(per, np.full(n_snr, C * ens))
        for r in range(world):
            si, ci, e = sharding.decode(sharding.frame_ids(n_snr, C, ens, (r, world)), C, ens)
            assert np.array_equal(np.bincount(si, minlength=n_snr), sharding.frames_per_snr(n_snr, C, ens, (r, world)))
            assert ci.max() < C and e.max() < ens
    bt, st = sharding.totals(n_snr, C, ens, 256, 16, 4)
    assert np.array_equal(st, np.full(n_snr, C * ens * 256 * 15)) and np.array_equal(bt, st * 4)


def test_mex_gateway_compiles():
    if not shutil.which("g++"):
        pytest.skip("no g++")
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-DWOFDM_MEX_STUB",
                    os.path.join(REPO, "mex", "wofdm_mex.cpp")], check=True)


def test_cpu_port_statistics():
    """The numba port timed by bench.py reproduces the oracle's SER statistically (same chain, own RNG)."""
    from oracle import wofdm_cpu_port as P
    ch = O.synth_channels(2, 21, seed=1)
    snr = np.array([5.0, 25.0])
    t = P.build_task("wtx", 256, 16, 8, 0, 8, 4, ch, 6, snr, 3)
    errs = P.run_task(t)
    ser_port = errs / (2 * 6 * 256 * 7)
    p = O.system_params("wtx", 256, 16, 8, 0, S=8)
    np.random.seed(5)
    ser = O.ser_sweep_replay(p, [(O.rc_window_tx(p), O.rc_window_rx(p))], ch, 6, snr, dense=False)[0]
    assert np.all(np.abs(ser_port - ser) < 4 * np.sqrt(ser * (1 - ser) / (2 * 6 * 256 * 7)) + 0.01)
