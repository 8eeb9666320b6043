"""World-size-2 (and 3) runs of the sharding logic on CPU with the gloo backend.

The per-rank evaluation is the C oracle here (no GPU in this container); what is under test is the host logic of
r2p2_b200/distributed.py: shard bounds, rank-ordered fitness all-gather, best-genome broadcast -- the result must be
bit-identical to the single-process evaluation of the whole population."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from r2p2_b200.distributed import ShardedEvaluator, broadcast_best, shard_bounds


def test_shard_bounds_cover_population():
    for P in (1, 7, 1024, 65536, 1000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(P, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == P
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _oracle_eval(genomes):
    from oracle.oracle import Oracle
    rob, kw = common.ROBOTS["robot-neuro.json"], common.oracle_robot_kwargs("robot-neuro.json")
    return Oracle(port=True).run(common.load_stage("track_2.png"), ctrl="neuro", genomes=genomes, x0=rob["x"], y0=rob["y"], theta0=0.0,
                                 T=60, layers=[5, 10, 7, 4, 2], **kw)["fitness"]


def _worker(rank, world, port, P, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genomes = np.random.RandomState(99).uniform(-1, 1, (P, 156))          # every rank holds the generation (replicated EA state)
    fit = ShardedEvaluator(_oracle_eval, device="cpu").evaluate(genomes)
    lo, hi = shard_bounds(P, world, rank)
    best, bf, bg = broadcast_best(torch.from_numpy(genomes[lo:hi].copy()), torch.from_numpy(fit), P)
    q.put((rank, fit, best, bf, bg.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,P", [(2, 48), (3, 50)])
def test_sharded_evaluation_equals_single_process(world, P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    genomes = np.random.RandomState(99).uniform(-1, 1, (P, 156))
    ref = _oracle_eval(genomes)
    for rank, fit, best, bf, bg in res:
        assert np.array_equal(fit, ref), rank                                 # bit-identical for every rank
        assert best == int(np.argmax(ref)) and bf == ref[best]
        assert np.array_equal(bg, genomes[best])
