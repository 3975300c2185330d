"""Multi-GPU use of the path: one process per GPU (torch.distributed).

Sampling needs no communication.  Streams are independent (ref: prng.hpp:11-14),
so each rank owns a contiguous block of streams and draws on its own device:

  * "long_jump" sharding (the reference's distributed-state scheme,
    R/random.R:161-171, R/sample.R:464-467): rank r's generator is the seed
    state long-jumped r times (2^192 steps apart), then a jump ladder inside
    the rank.  Benchmark default.
  * "offset" sharding: rank r holds streams [r*S/W, (r+1)*S/W) of the single
    reference generator prng(S, seed) (prng.hpp:38-53), bit-identical to
    monty_rng_create(n_streams = S) on one device; concatenating the ranks'
    monty_rng_state() in rank order gives the reference's state bytes.

The only collective on the path is the sum of log-likelihoods (SURVEY config
C5): each rank reduces its observations to one double on the device (fused in
the density kernel) and the 8-byte partials are all-reduced -- NCCL over
NVLink on GPUs, gloo in the CPU tests of this host logic.
"""
import ctypes as C

from . import _lib
from ._lib import DENS, DENS_N_PARAMS, check, lib
from .random import monty_rng_create


def shard_bounds(n_total, rank, world):
    """Contiguous, balanced partition: the first n_total % world ranks get one extra."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def create_sharded_rng(n_streams_total, seed, rank, world, mode="long_jump", device=None,
                       real_type="f64"):
    """This rank's part of an n_streams_total-stream generator."""
    lo, hi = shard_bounds(n_streams_total, rank, world)
    device = rank if device is None else device
    if mode == "long_jump":
        return monty_rng_create(hi - lo, seed, device=device, long_jumps=rank, real_type=real_type)
    if mode == "offset":
        return monty_rng_create(hi - lo, seed, device=device, stream_offset=lo, real_type=real_type)
    raise ValueError("mode must be 'long_jump' or 'offset'")


def density_sum_dev(name, x, params, log=True, real_type="f64", out=None):
    """Local part: sum of (log-)densities over this rank's observations,
    computed and reduced on the device.  x / params are torch CUDA tensors
    (int32 or float64, as the reference's cpp11::integers / doubles).  Returns
    a 1-element float64 CUDA tensor (no host synchronisation)."""
    import torch
    d = DENS[name]
    dev = x.device
    n = x.numel()
    ptrs = (C.c_void_p * 4)()
    isi = (C.c_int * 4)()
    for k, p in enumerate(params):
        if p.numel() < n:
            raise ValueError("density arguments must all have the length of x (no recycling)")
        ptrs[k] = p.data_ptr()
        isi[k] = 1 if p.dtype == torch.int32 else 0
    if len(params) != DENS_N_PARAMS[d]:
        raise ValueError("wrong number of parameters for %s" % name)
    total = torch.zeros(1, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(lib.mb200_density_dev(d, 1 if real_type == "f32" else 0, n,
                                C.c_void_p(x.data_ptr()), 1 if x.dtype == torch.int32 else 0,
                                ptrs, isi, 1 if log else 0,
                                C.c_void_p(out.data_ptr()) if out is not None else None,
                                C.c_void_p(total.data_ptr()), dev.index or 0, C.c_void_p(stream)))
    return total


def allreduce_sum(partial, group=None):
    """Sum a per-rank partial (tensor of doubles) over all ranks, in place."""
    import torch.distributed as td
    if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
        td.all_reduce(partial, op=td.ReduceOp.SUM, group=group)
    return partial


def log_likelihood(name, x, params, real_type="f64", group=None):
    """Global log-likelihood of observations sharded over the ranks."""
    return allreduce_sum(density_sum_dev(name, x, params, True, real_type), group)
