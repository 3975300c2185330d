#!/usr/bin/env python
"""bench.py -- throughput of the sampling hot path (draws/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--dist NAME] [--real f64|f32] [--streams S] [--draws D]
                    [--detail]

One "step" = one launch of the sampler over the whole workload: S streams x D
draws per stream (monty_random_n_<dist>(D, ..., state), ref: src/random.cpp:
211-338), output written stream-major to HBM.  The default workload is
BASELINE.json configs[1]: normal draws, ziggurat, 2^20 streams x 1000 draws,
double.  Prints ONE JSON line (see the contract in the task statement):

  value      whole-job Gdraws/s, inputs and outputs resident in HBM
  e2e        the same metric through the host-buffer C-ABI call
             (mb200_rng_sample): parameters copied host->device and all draws
             copied device->host into pinned memory inside the timed region
  roofline   algorithmic HBM bytes per launch / CUDA-event kernel time
             against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline  the reference's own CPU code (oracle/_ref, kind "reference";
             or the C restatement, kind "port") on this box's host cores

--impl reference runs the reference's CPU implementation only (rank 0).
Inputs (16 B of scalars) are far smaller than L2; the 8.6 GB of output per
step is far larger than L2 (126 MB), so no L2 flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

DEFAULT_DIST = "normal_ziggurat"
N_PARAMS = {"real": 0, "uniform": 2, "exponential_rate": 1, "normal_ziggurat": 2,
            "normal_box_muller": 2, "normal_polar": 2, "binomial": 2, "poisson": 1,
            "gamma_scale": 2, "beta": 2, "negative_binomial_prob": 2,
            "hypergeometric": 3, "truncated_normal": 4}


def synthetic_params(dist, n, first_stream=0):
    """Seed-defined per-stream parameters, identical on CPU and GPU.
    Scalars where the config uses scalars (C1/C2), vectors for C3/C4."""
    from oracle_lib import counter_hash
    idx = np.arange(first_stream, first_stream + n)
    h = lambda k: counter_hash(idx, k)  # noqa: E731
    if dist == "real":
        return []
    if dist in ("normal_ziggurat", "normal_box_muller", "normal_polar"):
        return [np.array([0.0]), np.array([1.0])]
    if dist == "uniform":
        return [np.array([0.0]), np.array([1.0])]
    if dist == "exponential_rate":
        return [np.array([1.0])]
    if dist == "binomial":
        return [np.floor(10 ** (6 * h(0))), h(1)]
    if dist == "poisson":
        lam = 10 ** (-1 + 5 * h(2))
        return [np.where(idx % 128 == 0, 1e8 * (1 + h(0)), lam)]
    if dist == "gamma_scale":
        return [10 ** (-1 + 2.5 * h(0)), 0.5 + h(1)]
    if dist == "beta":
        return [0.2 + 20 * h(0), 0.2 + 20 * h(1)]
    if dist == "negative_binomial_prob":
        return [0.5 + 100 * h(0), 0.05 + 0.9 * h(1)]
    if dist == "hypergeometric":
        n1 = np.floor(2000 * h(0)); n2 = np.floor(2000 * h(1))
        return [n1, n2, np.floor((n1 + n2) * h(2))]
    if dist == "truncated_normal":
        mn = -3 + 5 * h(0)
        mx = mn + 0.1 + 4 * h(1)
        mx = np.where(idx % 8 == 1, np.inf, mx)
        mn = np.where(idx % 16 == 0, -np.inf, mn)
        return [np.zeros(n), np.ones(n), mn, mx]
    raise ValueError(dist)


def algorithmic_bytes(n_streams, n_draws, params, out_bytes):
    """per draw sizeof(out) + (2*32 + 8*P_vec) / D."""
    p_vec = sum(1 for p in params if p.size > 1)
    return n_streams * (n_draws * out_bytes + 64 + 8 * p_vec)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(dist, real, n_streams, n_draws, steps, warmup, threads=None):
    """Time the reference's CPU implementation (all host threads, OpenMP static
    over streams as in inst/random/openmp/rnguse.cpp:19-24)."""
    from oracle_lib import Oracle, have_reference
    kind = "reference" if have_reference() else "port"
    orc = Oracle(kind)
    # all host cores this process may use (torchrun sets OMP_NUM_THREADS=1, so ask the OS)
    threads = threads or len(os.sched_getaffinity(0))
    st = orc.seed_states(42, n_streams)
    params = synthetic_params(dist, n_streams)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        orc.sample(dist, st, n_draws, params, real, n_threads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return kind, threads, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a bounded sample of the same workload: S/64 streams, same draws per stream
    n_streams = max(1024, args.streams // 64)
    kind, threads, times = cpu_reference_run(args.dist, args.real, n_streams, args.draws,
                                             args.steps, args.warmup)
    t = float(np.mean(times))
    value = n_streams * args.draws / t / 1e9
    sample = "%d streams x %d draws per step (1/%d of the workload's streams), %d threads" % (
        n_streams, args.draws, args.streams // n_streams, threads)
    line = {
        "impl": "reference", "metric": "Gdraws/s (%s, %s)" % (args.dist, args.real),
        "value": value, "unit": "Gdraws/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.real, "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "Gdraws/s", "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "Gdraws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "monty_random_n_%s: %d streams x %d draws/stream, real_type %s, seed 42, "
                        "long_jump-separated shard per GPU" % (args.dist, args.streams, args.draws, args.real),
            "streams_per_gpu": args.streams, "draws_per_stream": args.draws,
            "l2": "per-step output (%.2f GB) >> L2, no flush needed" % (
                args.streams * args.draws * (4 if args.real == "f32" and args.out_f32 else 8) / 1e9)}


def time_device(mb, torch, rng, dist, real, n_draws, params_dev, out, steps, warmup, barrier):
    stream = torch.cuda.ExternalStream(rng.cuda_stream(), device=out.device)
    for _ in range(warmup):
        rng.sample_dev(dist, n_draws, params_dev, out, real)
    rng.sync()
    barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    for a, b in evs:
        a.record(stream)
        rng.sample_dev(dist, n_draws, params_dev, out, real)
        b.record(stream)
    rng.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = evs[0][0].elapsed_time(evs[-1][1])
    return total_ms, kernel_ms, wall


def density_inputs(torch, name, n, first, dev):
    """SURVEY 8(d) C5 inputs, generated on the device from the same counter hash."""
    idx = torch.arange(first, first + n, dtype=torch.int64, device=dev)

    def h(k):
        m64 = (1 << 64) - 1
        # splitmix64 on int64 with wrap-around (two's complement == mod 2^64)
        def c(v):
            v &= m64
            return v - (1 << 64) if v >= (1 << 63) else v
        z = (idx * 8 + k) * 2 + 1
        z = z * c(0x9E3779B97F4A7C15)
        z = z ^ c(0xC0FFEE)
        z = z + c(0x9E3779B97F4A7C15)
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * c(0xBF58476D1CE4E5B9)
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * c(0x94D049BB133111EB)
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        top = (z >> 11) & ((1 << 53) - 1)
        return top.to(torch.float64) * 1.1102230246251565e-16 + 1.1102230246251565e-16 / 2.0
    x = torch.floor(200 * h(0) ** 2).to(torch.int32)
    if name == "poisson":
        return x, [0.5 + 100 * h(1)]
    if name == "negative_binomial_mu":
        return x, [0.5 + 50 * h(1), 0.1 + 100 * h(2)]
    size = (x + torch.floor(100 * h(1)).to(torch.int32)).to(torch.int32)
    return x, [size, 0.05 + 0.9 * h(2), 0.01 + 0.49 * h(3)]


def run_density(args, mb, torch, dev, rank, world, local_rank, dist_on, barrier):
    from monty_b200.distributed import density_sum_dev, allreduce_sum
    n = args.obs
    x, pars = density_inputs(torch, args.density, n, rank * n, dev)
    bytes_per_obs = x.element_size() + sum(p.element_size() for p in pars)
    launches0 = mb._lib.lib.mb200_launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        allreduce_sum(density_sum_dev(args.density, x, pars))
    torch.cuda.synchronize()
    barrier()
    evs = []
    total = None
    for _ in range(args.steps):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        part = density_sum_dev(args.density, x, pars)
        b.record()
        total = allreduce_sum(part)
        c.record()
        evs.append((a, b, c))
    torch.cuda.synchronize()
    barrier()
    clock_info = clocks.stop()
    launches = mb._lib.lib.mb200_launch_count() - launches0 - 2 * max(args.warmup, 3)
    k_ms = float(np.mean([a.elapsed_time(b) for a, b, c in evs]))
    tot_ms = evs[0][0].elapsed_time(evs[-1][2])
    t = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
    if dist_on:
        import torch.distributed as td
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    peak, peak_src = measured_peak()
    achieved = n * bytes_per_obs / (k_ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "Gobs/s (log-density %s, f64, fused sum + all-reduce)" % args.density,
            "value": world * n / (ms_per_step * 1e-3) / 1e9, "unit": "Gobs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "density::%s log-likelihood, %d observations per GPU, reduce-only"
                                   % (args.density, n), "l2": "inputs %.1f GB >> L2" % (n * bytes_per_obs / 1e9)},
            "clocks": clock_info, "gpu_launches": int(launches), "log_likelihood": float(total.item()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "density_kernel<%s>" % args.density, "kernel_ms": k_ms,
                         "note": "lgamma-heavy: expected FP64-pipe bound, not HBM bound"}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dist", default=DEFAULT_DIST)
    ap.add_argument("--real", default="f64", choices=["f64", "f32"])
    ap.add_argument("--streams", type=int, default=1 << 20)
    ap.add_argument("--draws", type=int, default=1000)
    ap.add_argument("--out-f32", dest="out_f32", action="store_true",
                    help="float32 device output (only with --real f32)")
    ap.add_argument("--detail", action="store_true", help="also time the other distributions (default at 1 GPU)")
    ap.add_argument("--no-detail", dest="no_detail", action="store_true",
                    help="skip the per-distribution table of the default 1-GPU run")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="sample", choices=["sample", "density"],
                    help="density: config C5, fused log-likelihood + all-reduce")
    ap.add_argument("--density", default="poisson",
                    choices=["poisson", "negative_binomial_mu", "beta_binomial_prob"])
    ap.add_argument("--obs", type=int, default=125_000_000, help="observations per GPU (density)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import monty_b200 as mb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=dev)
        barrier = lambda: td.barrier()  # noqa: E731
    else:
        barrier = lambda: None  # noqa: E731

    if args.workload == "density":
        run_density(args, mb, torch, dev, rank, world, local_rank, dist_on, barrier)
        if dist_on:
            td.barrier()
            td.destroy_process_group()
        return

    S, D = args.streams, args.draws
    # shard: GPU g holds S streams of the generator seed(42) long-jumped g times
    rng = mb.monty_rng_create(S, 42, device=local_rank, long_jumps=rank, real_type=args.real)
    params = synthetic_params(args.dist, S, first_stream=rank * S)
    params_dev = [torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in params]
    out_dtype = torch.float32 if (args.real == "f32" and args.out_f32) else torch.float64
    out = torch.empty((S, D), dtype=out_dtype, device=dev)

    launches0 = mb._lib.lib.mb200_launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    total_ms, kernel_ms, _ = time_device(mb, torch, rng, args.dist, args.real, D, params_dev, out,
                                         args.steps, args.warmup, barrier)
    clock_info = clocks.stop()
    launches = mb._lib.lib.mb200_launch_count() - launches0 - args.warmup

    t_local = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist_on:
        td.all_reduce(t_local, op=td.ReduceOp.MAX)
    t_ms = float(t_local.item())
    ms_per_step = t_ms / args.steps
    value = world * S * D / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant (only) kernel, per launch
    peak, peak_src = measured_peak()
    alg_bytes = algorithmic_bytes(S, D, params, out.element_size())
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel": ("ziggurat_kernel<%s>" % args.real) if (args.dist == "normal_ziggurat" and D >= 2)
                else "sample_kernel<%s,%s>" % (args.dist, args.real),
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as fh:
                tr = json.load(fh).get("%s_%s" % (args.dist, args.real))
            if tr:
                roofline["traffic"] = tr
        except Exception:
            pass

    # end to end through the host-buffer C-ABI call (what the R glue calls)
    e2e = None
    if rank == 0 or dist_on:
        try:
            host_out = torch.empty((S, D), dtype=torch.float64, pin_memory=True)
            pinned = True
        except RuntimeError:                     # many ranks on one host can exhaust lockable memory
            host_out = torch.empty((S, D), dtype=torch.float64)
            pinned = False
        host_np = host_out.numpy()
        import ctypes as C
        arrs = [np.ascontiguousarray(p) for p in params]
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs] + [None] * (4 - len(arrs)))
        lens = (C.c_size_t * 4)(*[a.size for a in arrs] + [0] * (4 - len(arrs)))
        lib = mb._lib.lib
        did = mb._lib.DIST[args.dist]
        rt = 1 if args.real == "f32" else 0

        def e2e_step():
            mb._lib.check(lib.mb200_rng_sample(rng.handle, did, rt, D, ptrs, lens,
                                               C.c_void_p(host_np.ctypes.data)))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist_on:
            td.all_reduce(t_e, op=td.ReduceOp.MAX)
        dt = float(t_e.item())
        e2e = {"value": world * S * D / dt / 1e9, "unit": "Gdraws/s",
               "h2d_bytes_per_step": int(sum(a.nbytes for a in arrs)),
               "d2h_bytes_per_step": int(S * D * 8), "ms_per_step": dt * 1e3,
               "api": "mb200_rng_sample (host buffers, %s output)" % ("pinned" if pinned else "pageable")}
        del host_out

    detail = None
    if (args.detail or (world == 1 and not args.no_detail)) and rank == 0:
        detail = {}
        for name, s_d in [("real", (S, D)), ("uniform", (S, D)), ("exponential_rate", (S, D)),
                          ("normal_box_muller", (S, D)), ("normal_polar", (S, D)),
                          ("binomial", (1 << 22, 1)), ("poisson", (1 << 22, 1)),
                          ("binomial", (1 << 22, 16)), ("poisson", (1 << 22, 16)),
                          ("gamma_scale", (1 << 21, 16)), ("beta", (1 << 21, 16)),
                          ("negative_binomial_prob", (1 << 21, 16)), ("hypergeometric", (1 << 21, 16)),
                          ("truncated_normal", (1 << 21, 16))]:
            s2, d2 = s_d
            r2 = mb.monty_rng_create(s2, 42, device=local_rank, real_type=args.real)
            p2 = synthetic_params(name, s2)
            p2d = [torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in p2]
            o2 = torch.empty((s2, d2), dtype=torch.float64, device=dev)
            tm, km, _ = time_device(mb, torch, r2, name, args.real, d2, p2d, o2, 5, 3, lambda: None)
            kms = float(np.mean(km))
            ab = algorithmic_bytes(s2, d2, p2, 8)
            detail["%s[%dx%d]" % (name, s2, d2)] = {
                "gdraws_per_s": s2 * d2 / (kms * 1e-3) / 1e9, "kernel_ms": kms,
                "hbm_gbs": ab / (kms * 1e-3) / 1e9, "hbm_frac": ab / (kms * 1e-3) / 1e9 / peak}
            r2.close()
            del o2

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = max(1024, S // 8)       # ~10 core-seconds of CPU work over the three repeats
        kind, threads, times = cpu_reference_run(args.dist, args.real, n_cpu, D, 3, 1)
        tcpu = float(np.mean(times))
        kind1, _, times1 = cpu_reference_run(args.dist, args.real, n_cpu // 8, D, 2, 1, threads=1)
        cpu = {"value": n_cpu * D / tcpu / 1e9, "unit": "Gdraws/s", "cores": threads, "kind": kind,
               "sample": "%d streams x %d draws (1/%d of the workload), OpenMP static over streams"
                         % (n_cpu, D, S // n_cpu),
               "single_thread_value": (n_cpu // 8) * D / float(np.mean(times1)) / 1e9}

    if rank == 0:
        line = {
            "metric": "Gdraws/s (%s, %s)" % (args.dist, args.real), "value": value,
            "unit": "Gdraws/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.real, "data": "synthetic",
            "config": workload_config(args), "clocks": clock_info,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if detail:
            line["detail"] = detail
        print(json.dumps(line))
    if dist_on:
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
