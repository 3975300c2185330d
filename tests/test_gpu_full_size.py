"""The configured sizes of BASELINE.json, on the device.

  * C4: 2^24 streams sharded by long_jump over 8 GPUs = 2^21 streams per GPU,
    rank r's generator = seed long-jumped r times then a jump ladder
    (ref: R/random.R:161-171).  For r in {0, 3, 7} a full 2^21-stream shard is
    created and sampled on this GPU with the shard's own per-stream parameters,
    and 4096 streams spread over it are compared with the reference run on the
    oracle's long-jumped generator: initial states, draws (bit-identical for the
    integer-valued samplers and for gamma with shape >= 1, see
    test_gpu_samplers.py), final states.
  * C5: 1.25e8 observations per GPU (1e9 over 8): a 1e5 sample of the
    per-observation log-densities against the reference (bit-identical: the
    device's log and lgamma return glibc's bits), and the device-side
    sum against the reference's sum of ALL 1.25e8 values within 1e-12 sqrt(N)
    relative.
  * the 2-rank NCCL path of monty_b200.distributed.log_likelihood with the GPU
    reduction (needs 2 GPUs: skipped on a 1-GPU box; run with
    `gpurun --gpus 2 -- python -m pytest tests/test_gpu_full_size.py -m gpu -k nccl`).
"""
import os
import sys

import numpy as np
import pytest

from cases import INTEGER_VALUED
from gpu_util import states_of, ulp_distance

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def mb():
    import monty_b200
    return monty_b200


_SHARD_CACHE = {}


def reference_shard_states(oracle, rank, pick):
    """States of streams `pick` of rank `rank`'s shard on the oracle: seed, `rank`
    long jumps, then sequential jumps (prng.hpp:38-53) -- 2^21 of them, so cached."""
    if rank not in _SHARD_CACHE:
        base = oracle.seed_states(42, 1)
        oracle.jump(base, rank, long=True)
        st = np.empty((pick.size, 4), dtype=np.uint64)
        cur, at = base.copy(), 0
        for i, k in enumerate(pick):                 # stream k of the shard = base jumped k times
            oracle.jump(cur, int(k - at))
            at = int(k)
            st[i] = cur[0]
        _SHARD_CACHE[rank] = st
    return _SHARD_CACHE[rank].copy()


@pytest.mark.parametrize("rank", [0, 3, 7])
@pytest.mark.parametrize("name", ["gamma_scale", "beta", "negative_binomial_prob", "hypergeometric",
                                  "truncated_normal"])
def test_c4_long_jump_shard_at_size(mb, oracle, name, rank):
    import torch
    from monty_b200.distributed import create_sharded_rng, shard_bounds
    from monty_b200.synthetic import sampler_params
    total, world, draws = 1 << 24, 8, 16
    lo, hi = shard_bounds(total, rank, world)
    S = hi - lo
    assert S == 1 << 21
    rng = create_sharded_rng(total, 42, rank, world, mode="long_jump", device=0)
    first = states_of(mb, rng)
    pick = np.linspace(0, S - 1, 4096).astype(np.int64)
    st = reference_shard_states(oracle, rank, pick)
    assert np.array_equal(first[pick], st), "shard states differ from the long-jumped reference generator"
    dev = torch.device("cuda", 0)
    pars = sampler_params(name, S, first_stream=lo)
    out = rng.sample_dev(name, draws, [torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in pars])
    rng.sync()
    last = states_of(mb, rng)
    psel = [p[pick] if p.size == S else p for p in pars]
    exp = oracle.sample(name, st, draws, psel)
    got = out[torch.from_numpy(pick).to(dev)].cpu().numpy()
    assert np.array_equal(last[pick], st), "final generator states differ"
    if name in INTEGER_VALUED or name == "truncated_normal":
        assert np.array_equal(got, exp)
    else:
        assert ulp_distance(got, exp).max() <= 4
        big = psel[0] >= 1 if name == "gamma_scale" else (psel[0] >= 1) & (psel[1] >= 1)
        assert np.array_equal(got[big], exp[big])          # no pow(): bit-identical


@pytest.mark.parametrize("name", ["poisson", "negative_binomial_mu", "beta_binomial_prob"])
def test_c5_at_size(mb, oracle, name):
    import torch
    from monty_b200.distributed import density_sum_dev
    from monty_b200.synthetic import density_inputs_torch, density_inputs_numpy
    n, rank = 125_000_000, 5                      # rank 5's slice of the 1e9 observations
    dev = torch.device("cuda", 0)
    x, pars = density_inputs_torch(name, n, rank * n, dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    total = density_sum_dev(name, x, pars, out=out)
    reduce_only = density_sum_dev(name, x, pars)
    torch.cuda.synchronize()
    assert total.item() == reduce_only.item()      # the sum does not depend on whether values are written
    # a 1e5 sample of the per-observation values
    idx = np.linspace(0, n - 1, 100_000).astype(np.int64)
    sel = torch.from_numpy(idx).to(dev)
    xs = x[sel].cpu().numpy()
    ps = [p[sel].cpu().numpy() for p in pars]
    exp = oracle.density(name, xs, ps, True)
    got = out[sel].cpu().numpy()
    assert np.array_equal(got, exp), "the three C5 log-densities are bit-identical to the reference"
    # the whole slice: host inputs from the same hash, reference evaluated on all host cores
    xh, ph = density_inputs_numpy(name, n, rank * n)
    assert np.array_equal(xh[idx], xs)
    ref = oracle.density(name, xh, ph, True, n_threads=len(os.sched_getaffinity(0)))
    ref_sum = float(np.sum(ref.astype(np.longdouble)))
    assert abs(total.item() - ref_sum) <= 1e-12 * np.sqrt(n) * abs(ref_sum), (total.item(), ref_sum)
    # and the device's values add up to the device's sum
    assert abs(out.sum().item() - total.item()) <= 1e-12 * abs(ref_sum)


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from monty_b200.distributed import log_likelihood, shard_bounds
    from monty_b200.synthetic import density_inputs_torch
    n = 4_000_000
    lo, hi = shard_bounds(n, rank, world)
    res = {}
    for name in ("poisson", "negative_binomial_mu", "beta_binomial_prob"):
        x, pars = density_inputs_torch(name, hi - lo, lo, dev)
        res[name] = float(log_likelihood(name, x, pars).item())     # device reduction + NCCL all-reduce
    q.put((rank, res))
    td.barrier()
    td.destroy_process_group()


def test_nccl_two_rank_log_likelihood_uses_the_gpu_reduction(oracle):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from monty_b200.synthetic import density_inputs_numpy
    world, port = 2, 29600 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n = 4_000_000
    for name in ("poisson", "negative_binomial_mu", "beta_binomial_prob"):
        assert res[0][name] == res[1][name]                       # every rank holds the same total
        x, pars = density_inputs_numpy(name, n)
        ref = float(np.sum(oracle.density(name, x, pars, True, n_threads=len(os.sched_getaffinity(0))).astype(np.longdouble)))
        assert abs(res[0][name] - ref) <= 1e-12 * np.sqrt(n) * abs(ref)
