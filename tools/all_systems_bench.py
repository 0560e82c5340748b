"""Device-time table of K1 over the reference's seven systems (development aid, not the bench).

One process: for every (system, N) a plan over 250 channels x 30 SNR points, RC windows, three warm-up launches, five timed
ones (CUDA events on the launching stream).  `python tools/all_systems_bench.py [ens256 [ens512 [ens1024]]]`."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wofdm_b200 as W
from wofdm_b200 import capi

SYSTEMS = ["CP", "wtx", "CPwtx", "wrx", "CPwrx", "WOLA", "CPW"]


def one(h, st, name, N, ens, L=21):
    scale = N // 256
    cp, ttx, trx = 16 * scale, 8 * scale, 10 * scale
    if name in ("CP", "wrx", "CPwrx"): ttx = 0
    if name in ("CP", "wtx", "CPwtx"): trx = 0
    s = W.params_from_name(name, N, cp, ttx, trx, bits=4 if N == 256 else 6, S=16)
    vt, vr = capi.rc_window_tx(s), capi.rc_window_rx(s)
    rng = np.random.default_rng(0)
    C = 250
    chan = (rng.standard_normal((L, C)) + 1j * rng.standard_normal((L, C))) * np.exp(-np.arange(L) / 4)[:, None]
    snr = np.linspace(-20, 50, 30)
    plan = h.ber_plan(s, vt, vr, chan, snr)
    for _ in range(3):
        plan.launch(ens, seed=1, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    e0.record()
    for k in range(K):
        plan.launch(ens, seed=2 + k, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    frames = 30 * C * ens
    rec = {"system": name, "N": N, "L": L, "kernel": plan.kernel, "frames": frames, "ms_per_launch": ms, "symbols_per_s": frames * 16 / ms * 1e3}
    plan.read()
    plan.close()
    return rec


def main():
    ens = [int(a) for a in sys.argv[1:4]] + [27, 8, 4][len(sys.argv[1:4]):]
    h = W.Handle([0])
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    for N, e in zip((256, 512, 1024), ens):
        for name in SYSTEMS:
            try:
                print(json.dumps(one(h, st, name, N, e)), flush=True)
            except Exception as ex:      # a shape the tensor-core kernels refuse would show here, not abort the table
                print(json.dumps({"system": name, "N": N, "error": str(ex)}), flush=True)
    for N, e in ((256, ens[0]), (1024, ens[2])):
        print(json.dumps(one(h, st, "WOLA", N, e, L=84)), flush=True)
    h.close()


main()
