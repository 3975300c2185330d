"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes.

The kernels need a GPU, so here the per-rank partial log-likelihoods come from
the oracle (test infrastructure standing in for the device reduction); what is
under test is the sharding arithmetic and the collective wiring of
monty_b200/distributed.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shard_bounds_partition():
    sys.path.insert(0, ROOT)
    from monty_b200.distributed import shard_bounds
    for n in (1, 7, 8, 1000, 2 ** 24, 10 ** 9):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from monty_b200.distributed import allreduce_sum, shard_bounds
    from oracle_lib import Oracle
    from cases import density_cases
    orc = Oracle("port")
    n = 40000
    x, pars = density_cases(n)["poisson"]
    lo, hi = shard_bounds(n, rank, world)
    part = orc.density("poisson", x[lo:hi], [pars[0][lo:hi]], True).sum()
    total = allreduce_sum(torch.tensor([part], dtype=torch.float64))
    # stream sharding: rank r's block of the single reference generator equals
    # the slice of the whole (offset mode), and the long-jump ladder equals
    # long_jump applied r times (long_jump mode)
    S = 64
    slo, shi = shard_bounds(S, rank, world)
    whole = orc.seed_states(42, S)
    base = orc.seed_states(42, 1)
    orc.jump(base, rank, long=True)
    ladder = orc.seed_states(0, shi - slo, base.ravel())
    q.put((rank, float(total.item()), whole[slo:shi].tobytes(), ladder.tobytes()))
    td.barrier()
    td.destroy_process_group()


def test_two_rank_gloo_allreduce_and_sharding():
    sys.path.insert(0, HERE)
    from oracle_lib import Oracle, build_oracles
    from cases import density_cases
    build_oracles()
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = Oracle("port")
    x, pars = density_cases(40000)["poisson"]
    ref = orc.density("poisson", x, pars, True).sum()
    for _, total, _, _ in res:
        assert abs(total - ref) <= 1e-12 * abs(ref)
    assert res[0][1] == res[1][1]
    whole = orc.seed_states(42, 64)
    assert b"".join(r[2] for r in res) == whole.tobytes()
    # long-jump shards are disjoint from each other and from the jump ladder
    assert res[0][3] == whole[:32].tobytes()
    assert res[1][3] != whole[32:].tobytes()
