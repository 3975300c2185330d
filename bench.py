#!/usr/bin/env python
"""bench.py -- throughput of the sampling hot path (draws/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--dist NAME] [--real f64|f32] [--streams-total S | --streams S_PER_GPU]
                    [--draws D] [--no-detail] [--no-density] [--no-cpu]

One "step" = one launch of the sampler over the whole workload: S streams x D
draws per stream (monty_random_n_<dist>(D, ..., state), ref: src/random.cpp:
211-338), output written stream-major to HBM.

Default workload = BASELINE.json's metric: normal draws (ziggurat, double) over
2^24 streams IN TOTAL, sharded over the N GPUs by long_jump (each rank holds
2^24 / N streams: strong scaling), D = 128 draws per stream per launch
(17.2 GB of output in total; D = 1000 would be 134 GB on one GPU and has no
host-side end-to-end leg).  `--streams S` switches to S streams PER GPU (weak
scaling); BASELINE configs[1] (2^20 x 1000) is `--streams 1048576 --draws 1000`
and is also timed as the first row of `detail`.

Prints ONE JSON line (the contract in the task statement):

  value        whole-job Gdraws/s, inputs and outputs resident in HBM
  e2e          the same metric through the host-buffer C-ABI call
               (mb200_rng_sample): parameters copied host->device and all draws
               copied device->host inside the timed region
  roofline     algorithmic HBM bytes per launch / CUDA-event kernel time against
               MEASURED_PEAKS.json's hbm_gbs, plus the pipe fractions of the
               kernel (instruction counts per draw from profiles/inst_counts.json,
               one ncu pass; peaks from profiles/pipe_peaks.json)
  cpu_baseline the reference's own CPU code on this box's host cores
  density      config C5: fused log-likelihood (Poisson, negative binomial,
               beta-binomial) of 1e9 observations in total + all-reduce, every N
  detail       (1 GPU) the other samplers at the C2 / C3 / C4 shapes with the
               bound that applies to each, and end-to-end legs of C3 / C4

--impl reference runs the reference's CPU implementation only (rank 0).
Inputs are a few bytes to 0.4 GB, outputs 2-17 GB per step, far larger than L2
(126 MB), so no L2 flush is needed between steps.
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

import numpy as np  # noqa: E402

DEFAULT_DIST = "normal_ziggurat"
STREAMS_TOTAL = 1 << 24
DRAWS_STRONG = 128
DRAWS_WEAK = 1000


def synthetic_params(dist, n, first_stream=0, per_stream_normal=False):
    from monty_b200.synthetic import sampler_params
    return sampler_params(dist, n, first_stream, per_stream_normal)


def algorithmic_bytes(n_streams, n_draws, params, out_bytes):
    """per draw sizeof(out) + (2*32 + 8*P_vec) / D."""
    p_vec = sum(1 for p in params if p.size > 1)
    return n_streams * (n_draws * out_bytes + 64 + 8 * p_vec)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region: NVML polled every 2 ms (the
    timed region of the default run is ~100 ms, one nvidia-smi process start), nvidia-smi
    as the fallback when the NVML binding is missing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.sm, self.mx, self.reasons, self.n = [], [], set(), 0
        self._stop_evt = threading.Event()
        self.source = "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self._nv = None
            self.source = "nvidia-smi"

    def _poll_nvml(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)
        self.n += 1

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                              "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        if r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        if r[1].replace(".", "").isdigit():
            self.mx.append(float(r[1]))
        for (name, _), v in zip(self.REASONS, r[2:6]):
            if v.lower().startswith("active"):
                self.reasons.add(name)
        self.n += 1

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nv is not None:
                    self._poll_nvml()
                else:
                    self._poll_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.002 if self._nv is not None else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": self.n, "source": self.source}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _load_json(name):
    p = os.path.join(ROOT, "profiles", name)
    try:
        with open(p) as fh:
            return json.load(fh)
    except Exception:
        return {}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def bounds(dist, n_draws, draws_per_s, hbm_gbs, hbm_peak, per_stream=False):
    """Which resource bounds the kernel: achieved fraction of the measured HBM
    peak and of each pipe's measured peak (profiles/pipe_peaks.json), from the
    kernel's warp-instruction counts per draw (profiles/inst_counts.json: one
    `ncu --metrics smsp__inst_executed*` pass per sampler and shape class,
    tools/ncu_counts.sh).  `bound` names the largest fraction."""
    fr = {"hbm": hbm_gbs / hbm_peak}
    counts = _load_json("inst_counts.json")
    peaks = _load_json("pipe_peaks.json")
    key = "%s@%d" % (dist, n_draws) + ("/per_stream" if per_stream else "")
    c = counts.get(key)
    if c and peaks:
        per = {"fp64": ("fp64_per_draw", "dfma_thread_inst_per_s"), "alu": ("alu_per_draw", "alu_thread_inst_per_s"),
               "fma": ("fma_per_draw", "ffma_thread_inst_per_s"), "issue": ("inst_per_draw", "issue_thread_inst_per_s")}
        for name, (ck, pk) in per.items():
            if ck in c and pk in peaks:
                fr[name] = c[ck] * 32.0 * draws_per_s / peaks[pk]
    bound = max(fr, key=lambda k: fr[k])
    out = {"bound": bound, "frac": fr[bound]}
    out.update({k + "_frac": v for k, v in fr.items()})
    if c:
        out["inst_per_draw"] = c.get("inst_per_draw")
        out["fp64_inst_per_draw"] = c.get("fp64_per_draw")
        out["lanes_active"] = c.get("lanes_active")
    else:
        out["pipe_fractions"] = "no instruction counts for %s in profiles/inst_counts.json" % key
    return out


def cpu_reference_run(dist, real, n_streams, n_draws, steps, warmup, threads=None):
    """Time the reference's CPU implementation (all host threads, OpenMP static
    over streams as in inst/random/openmp/rnguse.cpp:19-24)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
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


def cpu_density_run(name, n_obs, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, have_reference
    from monty_b200.synthetic import density_inputs_numpy
    kind = "reference" if have_reference() else "port"
    orc = Oracle(kind)
    x, pars = density_inputs_numpy(name, n_obs)
    orc.density(name, x[:1000], [p[:1000] for p in pars], True)
    t0 = time.perf_counter()
    orc.density(name, x, pars, True)
    return kind, time.perf_counter() - t0


def workload(args, world):
    """(streams on this GPU's shard description, config dict)."""
    if args.streams:
        total = args.streams * world
        desc = "%d streams per GPU" % args.streams
        scaling = "weak"
    else:
        total = args.streams_total
        desc = "%d streams in total (%d per GPU)" % (total, total // world)
        scaling = "strong"
    out_bytes = 4 if (args.real == "f32" and args.out_f32) else 8
    cfg = {"workload": "monty_random_n_%s: %s x %d draws/stream, real_type %s, seed 42, long_jump-separated "
                       "shard per GPU" % (args.dist, desc, args.draws, args.real),
           "streams_total": total, "streams_per_gpu": total // world, "draws_per_stream": args.draws,
           "l2": "per-step output per GPU (%.2f GB) >> L2 (126 MB), no flush needed"
                 % (total // world * args.draws * out_bytes / 1e9)}
    return total, scaling, cfg


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    total, scaling, cfg = workload(args, world)
    # a bounded sample of the same workload: 1/64 of the streams, same draws per stream
    n_streams = max(1024, total // 64)
    kind, threads, times = cpu_reference_run(args.dist, args.real, n_streams, args.draws,
                                             args.steps, args.warmup)
    t = float(np.mean(times))
    value = n_streams * args.draws / t / 1e9
    sample = "%d streams x %d draws per step (1/%d of the workload's streams), %d threads" % (
        n_streams, args.draws, total // n_streams, threads)
    line = {
        "impl": "reference", "metric": "Gdraws/s (%s, %s)" % (args.dist, args.real),
        "value": value, "unit": "Gdraws/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": args.real, "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "Gdraws/s", "cores": threads, "cpu": cpu_model(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "Gdraws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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


def host_buffer(torch, shape, want_pinned=True):
    if want_pinned:
        try:
            return torch.empty(shape, dtype=torch.float64, pin_memory=True), True
        except RuntimeError:            # many ranks on one host can exhaust lockable memory
            pass
    return torch.empty(shape, dtype=torch.float64), False


def time_e2e(mb, torch, rng, dist, real, n_streams, n_draws, params, steps, barrier, pinned=True):
    """Seconds per call of mb200_rng_sample with HOST parameter and output buffers."""
    import ctypes as C
    host_out, is_pinned = host_buffer(torch, (n_streams, n_draws), pinned)
    host_np = host_out.numpy()
    arrs = [np.ascontiguousarray(p) for p in params]
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs] + [None] * (4 - len(arrs)))
    lens = (C.c_size_t * 4)(*[a.size for a in arrs] + [0] * (4 - len(arrs)))
    lib = mb._lib.lib
    did = mb._lib.DIST[dist]
    rt = 1 if real == "f32" else 0

    def step():
        mb._lib.check(lib.mb200_rng_sample(rng.handle, did, rt, n_draws, ptrs, lens,
                                           C.c_void_p(host_np.ctypes.data)))
    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    h2d = int(sum(a.nbytes for a in arrs))
    d2h = int(n_streams * n_draws * 8)
    del host_out
    return dt, is_pinned, h2d, d2h


def density_leg(args, mb, torch, dev, rank, world, dist_on, barrier, hbm_peak):
    """Config C5: per-rank fused log-likelihood + all-reduce of the 8-byte partials."""
    from monty_b200.distributed import density_sum_dev, allreduce_sum
    from monty_b200.synthetic import density_inputs_torch, DENSITY_BYTES_PER_OBS
    n_total = args.obs_total
    lo = rank * (n_total // world)
    n = n_total // world
    res = {}
    for name in ("poisson", "negative_binomial_mu", "beta_binomial_prob"):
        x, pars = density_inputs_torch(name, n, lo, dev)
        for _ in range(3):
            allreduce_sum(density_sum_dev(name, x, pars))
        torch.cuda.synchronize()
        barrier()
        steps = 10
        evs = []
        total = None
        for _ in range(steps):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            part = density_sum_dev(name, x, pars)
            b.record()
            total = allreduce_sum(part)
            c.record()
            evs.append((a, b, c))
        torch.cuda.synchronize()
        barrier()
        k_ms = float(np.mean([a.elapsed_time(b) for a, b, c in evs]))
        ar_us = float(np.mean([b.elapsed_time(c) for a, b, c in evs])) * 1e3
        tot_ms = evs[0][0].elapsed_time(evs[-1][2]) / steps
        t = torch.tensor([tot_ms, k_ms, ar_us], dtype=torch.float64, device=dev)
        if dist_on:
            import torch.distributed as td
            td.all_reduce(t, op=td.ReduceOp.MAX)
        tot_ms, k_ms, ar_us = (float(v) for v in t.tolist())
        bpo = DENSITY_BYTES_PER_OBS[name]
        gbs = n * bpo / (k_ms * 1e-3) / 1e9
        res[name] = {"gobs_per_s": world * n / (tot_ms * 1e-3) / 1e9, "ms_per_evaluation": tot_ms,
                     "kernel_ms": k_ms, "allreduce_us": ar_us if dist_on else 0.0,
                     "observations_total": world * n, "observations_per_gpu": n, "n_gpus": world,
                     "bytes_per_obs": bpo, "hbm_gbs_per_gpu": gbs, "hbm_frac": gbs / hbm_peak,
                     "log_likelihood": float(total.item())}
        del x, pars
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dist", default=DEFAULT_DIST)
    ap.add_argument("--real", default="f64", choices=["f64", "f32"])
    ap.add_argument("--streams-total", dest="streams_total", type=int, default=STREAMS_TOTAL,
                    help="streams over all GPUs (strong scaling; the default workload)")
    ap.add_argument("--streams", type=int, default=0, help="streams PER GPU (weak scaling) instead of --streams-total")
    ap.add_argument("--draws", type=int, default=0, help="draws per stream per launch (default 128; 1000 with --streams)")
    ap.add_argument("--per-stream-params", action="store_true", help="normal samplers: per-stream mean / sd vectors")
    ap.add_argument("--out-f32", dest="out_f32", action="store_true",
                    help="float32 device output (only with --real f32)")
    ap.add_argument("--detail", action="store_true", help="also time the other distributions (default at 1 GPU)")
    ap.add_argument("--no-detail", dest="no_detail", action="store_true",
                    help="skip the per-distribution table of the default 1-GPU run")
    ap.add_argument("--no-density", action="store_true", help="skip the config-C5 log-likelihood leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (development)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--obs-total", dest="obs_total", type=int, default=1_000_000_000,
                    help="observations over all GPUs in the density leg (config C5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if not args.draws:
        args.draws = DRAWS_WEAK if args.streams else DRAWS_STRONG

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import monty_b200 as mb
    from monty_b200.distributed import create_sharded_rng
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

    total, scaling, cfg = workload(args, world)
    D = args.draws
    # shard: rank r holds its block of streams of the generator seed(42) long-jumped r times
    # (R/random.R:161-171); no collective on the sampling path
    if args.streams:
        S = args.streams
        first = rank * S
        rng = mb.monty_rng_create(S, 42, device=local_rank, long_jumps=rank, real_type=args.real)
    else:
        from monty_b200.distributed import shard_bounds
        first, last = shard_bounds(total, rank, world)
        S = last - first
        rng = create_sharded_rng(total, 42, rank, world, mode="long_jump", device=local_rank, real_type=args.real)
    params = synthetic_params(args.dist, S, first_stream=first, per_stream_normal=args.per_stream_params)
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
    value = total * D / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant (only) kernel, per launch, on this rank's shard
    peak, peak_src = measured_peak()
    alg_bytes = algorithmic_bytes(S, D, params, out.element_size())
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    kernel_name = ("ziggurat_kernel<%s>" % args.real) if (args.dist == "normal_ziggurat" and D >= 2) \
        else "sample_kernel<%s,%s>" % (args.dist, args.real)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                "kernel": kernel_name, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "launch_shape": "%d streams x %d draws on one GPU" % (S, D)}
    tr = _load_json("traffic.json").get("%s_%s@%dx%d" % (args.dist, args.real, S, D))
    if tr:
        roofline["traffic"] = tr["bytes"]
        roofline["traffic_source"] = "profiles/traffic.json: %s (static file from one ncu --set full capture of this " \
                                     "launch shape, NOT measured in this run)" % tr["source"]
    roofline["pipes"] = bounds(args.dist, D, S * D / (k_ms * 1e-3), achieved, peak, args.per_stream_params)

    # end to end through the host-buffer C-ABI call (what the R glue calls)
    e2e = None
    if not args.no_e2e:
        dt, pinned, h2d, d2h = time_e2e(mb, torch, rng, args.dist, args.real, S, D, params, args.e2e_steps, barrier)
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        flags = [pinned]
        if dist_on:
            td.all_reduce(t_e, op=td.ReduceOp.MAX)
            flags = [None] * world
            td.all_gather_object(flags, pinned)
        dt = float(t_e.item())
        e2e = {"value": total * D / dt / 1e9, "unit": "Gdraws/s",
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "ms_per_step": dt * 1e3,
               "timed_calls": args.e2e_steps, "pinned_output_per_rank": flags,
               "api": "mb200_rng_sample (host parameter and output buffers), one call per rank per step"}
    del out

    # config C5 at every N: fused log-likelihood + all-reduce
    density = None
    if not args.no_density:
        density = density_leg(args, mb, torch, dev, rank, world, dist_on, barrier, peak)

    detail = None
    if (args.detail or (world == 1 and not args.no_detail)) and rank == 0:
        detail = {}
        c2 = (1 << 20, 1000)
        rows = [("normal_ziggurat", c2, False), ("normal_ziggurat", c2, True),
                ("real", c2, False), ("uniform", c2, False), ("exponential_rate", c2, False),
                ("normal_box_muller", c2, False), ("normal_polar", c2, False),
                ("binomial", (1 << 22, 1), False), ("poisson", (1 << 22, 1), False),
                ("binomial", (1 << 22, 16), False), ("poisson", (1 << 22, 16), False),
                ("gamma_scale", (1 << 21, 16), False), ("beta", (1 << 21, 16), False),
                ("negative_binomial_prob", (1 << 21, 16), False), ("hypergeometric", (1 << 21, 16), False),
                ("truncated_normal", (1 << 21, 16), False)]
        for name, (s2, d2), per_stream in rows:
            r2 = mb.monty_rng_create(s2, 42, device=local_rank, real_type=args.real)
            p2 = synthetic_params(name, s2, per_stream_normal=per_stream)
            p2d = [torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in p2]
            o2 = torch.empty((s2, d2), dtype=torch.float64, device=dev)
            tm, km, _ = time_device(mb, torch, r2, name, args.real, d2, p2d, o2, 5, 3, lambda: None)
            kms = float(np.mean(km))
            ab = algorithmic_bytes(s2, d2, p2, 8)
            gbs = ab / (kms * 1e-3) / 1e9
            row = {"gdraws_per_s": s2 * d2 / (kms * 1e-3) / 1e9, "kernel_ms": kms, "hbm_gbs": gbs}
            row.update(bounds(name, d2, s2 * d2 / (kms * 1e-3), gbs, peak, per_stream))
            if d2 <= 16 and name in ("binomial", "poisson", "gamma_scale", "hypergeometric") and not args.no_e2e:
                # the shapes the R API is called in a loop with: host vectors up, draws down
                for pin in (True, False):
                    dt, was, h2d, d2h = time_e2e(mb, torch, r2, name, args.real, s2, d2, p2, 10, lambda: None, pin)
                    row["e2e_%s" % ("pinned" if was else "pageable")] = {
                        "gdraws_per_s": s2 * d2 / dt / 1e9, "ms_per_call": dt * 1e3, "timed_calls": 10,
                        "h2d_bytes": h2d, "d2h_bytes": d2h}
            detail["%s[%dx%d]%s" % (name, s2, d2, " per-stream mean/sd" if per_stream else "")] = row
            r2.close()
            del o2

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = max(1024, total // 128)       # ~10 core-seconds of CPU work over the three repeats
        kind, threads, times = cpu_reference_run(args.dist, args.real, n_cpu, D, 3, 1)
        tcpu = float(np.mean(times))
        kind1, _, times1 = cpu_reference_run(args.dist, args.real, n_cpu // 8, D, 2, 1, threads=1)
        cpu = {"value": n_cpu * D / tcpu / 1e9, "unit": "Gdraws/s", "cores": threads, "cpu": cpu_model(), "kind": kind,
               "sample": "%d streams x %d draws (1/%d of the workload), OpenMP static over streams"
                         % (n_cpu, D, total // n_cpu),
               "single_thread_value": (n_cpu // 8) * D / float(np.mean(times1)) / 1e9}
        if density:
            for name in density:
                kd, td_s = cpu_density_run(name, 2_000_000)
                density[name]["cpu_gobs_per_s_single_thread"] = 2_000_000 / td_s / 1e9

    if rank == 0:
        line = {
            "metric": "Gdraws/s (%s, %s)" % (args.dist, args.real), "value": value,
            "unit": "Gdraws/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": args.real, "data": "synthetic",
            "config": cfg, "clocks": clock_info,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if density:
            line["density"] = density
        if detail:
            line["detail"] = detail
        print(json.dumps(line))
    if dist_on:
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
