"""World-size-2 gloo test (CPU) of the N > 1 host path: shard the frame ids, count per shard, one int64
all-reduce -> every rank holds the unsharded counters."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wofdm_b200 import sharding

N_SNR, C, ENS = 4, 5, 3


def fake_frame_errors(f):
    """deterministic stand-in for a frame's (bit_err, sym_err): a function of the GLOBAL frame id only"""
    f = np.asarray(f, dtype=np.int64)
    return (f * 2654435761 % 97), (f * 40503 % 13)


def counters_for(shard):
    ids = sharding.frame_ids(N_SNR, C, ENS, shard)
    si, _, _ = sharding.decode(ids, C, ENS)
    be, se = fake_frame_errors(ids)
    out = np.zeros((N_SNR, 2), dtype=np.int64)
    np.add.at(out[:, 0], si, be)
    np.add.at(out[:, 1], si, se)
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = sharding.allreduce_counters(torch.from_numpy(counters_for((rank, world))))
    tot = sharding.allreduce_counters(torch.from_numpy(np.stack(sharding.totals(N_SNR, C, ENS, 256, 16, 4, (rank, world)))))
    q.put((rank, t.numpy().copy(), tot.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_counter_allreduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = counters_for((0, 1))
    want_tot = np.stack(sharding.totals(N_SNR, C, ENS, 256, 16, 4))
    for _, t, tot in got:
        assert np.array_equal(t, want)
        assert np.array_equal(tot, want_tot)
