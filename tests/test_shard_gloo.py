"""Multi-rank host logic on CPU: world_size-2 gloo, the oracle standing in for the count
kernel (this file is test infrastructure; the product path runs the CUDA kernels per rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgcount_b200 import shard

from helpers import make_library, make_reads, oracle_library


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_covers_every_read_once():
    for world in (1, 2, 3, 8):
        for sizes in ([1000], [5, 0, 77], [10_000, 20_000, 30_000, 1], [123_457] * 8, [50_000_000]):
            plan = shard.plan_shards(sizes, world)
            for s, total in enumerate(sizes):
                mine = sorted((p.first_read, p.n_reads) for p in plan if p.sample == s)
                pos = 0
                for a, n in mine:
                    assert a == pos
                    pos += n
                assert pos == total
            assert all(0 <= p.rank < max(world, 1) for p in plan)
            if len(sizes) >= world:
                assert not shard.samples_spanning_ranks(plan)  # whole samples: no exchange
            else:
                assert all(p.first_read % 256 == 0 for p in plan)


def test_plan_balances_samples():
    plan = shard.plan_shards([100, 100, 100, 100, 400, 400, 50, 50], 4)
    load = [sum(p.n_reads for p in plan if p.rank == r) for r in range(4)]
    assert max(load) <= 450


def _oracle_state(guides, seqs, offset):
    from oracle import oracle as orc

    olib, _ = oracle_library(guides)
    c = orc.Counter.new(orc.Records.from_seqs(seqs), olib, orc.Permuter.new(olib), orc.Offset(*offset), None, True,
                        n_threads=1)
    return np.concatenate([c.counts_by_index().astype(np.int64), [c.total_reads(), c.matched_reads()]])


def _worker(rank, world, port, sizes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        guides = make_library(rng, 120, 20)
        samples = [make_reads(np.random.default_rng(100 + s), guides, n, 60, 7) for s, n in enumerate(sizes)]
        plan = shard.plan_shards(sizes, world, align=16)
        # offsets: decided by the owner of the sample's first shard, broadcast to the others
        offsets = []
        for s in range(len(sizes)):
            src = shard.owner_of(plan, s)
            local = (False, 7) if rank == src else (True, 999)  # a non-owner's guess must be overwritten
            offsets.append(shard.broadcast_offset(local, src))
        assert offsets == [(False, 7)] * len(sizes)

        def count_shard(sh):
            seqs = samples[sh.sample][sh.first_read:sh.first_read + sh.n_reads]
            return torch.from_numpy(_oracle_state(guides, seqs, offsets[sh.sample]))

        table = shard.count_samples(plan, len(sizes), len(guides), rank, count_shard)
        want = np.stack([_oracle_state(guides, samples[s], (False, 7)) for s in range(len(sizes))])
        assert np.array_equal(table.numpy(), want)
        assert table[:, len(guides)].tolist() == list(sizes)  # total_reads per sample
        if rank == 0:
            q.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes", [(700,), (300, 200, 100)])
def test_world_size_two_matches_single_process(sizes):
    """one sample read-sharded over 2 ranks (all-reduce does the sum) and 3 samples dealt to 2
    ranks (all-reduce only delivers columns): both equal the single-process table"""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() == "ok"
