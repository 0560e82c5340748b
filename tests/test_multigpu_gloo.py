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
    # interference: contiguous channel blocks, one all-gather of the per-channel rows
    Cc = 7
    lo, hi = sharding.channel_block(Cc, rank, world)
    P_local = np.arange(lo, hi, dtype=np.float64)[:, None] * np.ones((1, 4)) + 0.5
    P = sharding.allgather_channel_rows(P_local, Cc)
    q.put((rank, t.numpy().copy(), tot.numpy().copy(), P.numpy().copy()))
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
    for _, t, tot, P in got:
        assert np.array_equal(t, want)
        assert np.array_equal(tot, want_tot)
        assert np.array_equal(P, np.arange(7, dtype=np.float64)[:, None] * np.ones((1, 4)) + 0.5)


def test_channel_blocks_partition():
    for C in (1, 7, 250, 10000):
        for world in (1, 2, 3, 8):
            blocks = [sharding.channel_block(C, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == C
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
