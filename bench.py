#!/usr/bin/env python
"""bench.py -- the registration hot path on 1..N B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c5] [--impl ours|reference]

A "step" is one registration of one synthetic scan pair: ProbPointCloudRegistration constructor + align() to the
reference's own stopping rule (src/prob_point_cloud_registration.cc:15-158) -- octree build, then per outer
iteration radius search -> weights + normal equations per LM iteration -> pose update -> cloud move -> convergence
test, all on the device.  The default workload is BASELINE.json configs[2], the 1M-point pair the metric is quoted
on ("c3": -m 10 -r 0.5 -d 5).  With N > 1 every rank registers its own pair of the same shape (independent scan
pairs, data-parallel, no data-path collective): weak scaling.  That is the headline `value`.

The two configs BASELINE.json PARTITIONS across GPUs ride in the same JSON line as sub-records, measured at the same N:

  batch_c5    BASELINE configs[4]: 1024 independent 120k-point pairs (CLI defaults), dealt to the ranks in contiguous
              blocks (multi.deal_pairs) and run through ppcr_align_batch on every rank; pairs/s = 1024 / max-over-ranks
              time: STRONG scaling.  `single_gpu_pairs_per_s` is rank 0 alone on a share of the batch while the others idle.
  sharded_c4  BASELINE configs[3]: ONE 10M-point pair, source slices over the ranks, target replicated
              (multi.ShardedRegistration), the 24 moments exchanged peer to peer from inside the evaluation kernel; ms per
              registration, with `sharded_parity`: every rank holds a bit-identical pose history and it equals the
              single-GPU registration of the same pair (run on every rank beside it).

  value   correspondences/s with the clouds already resident in HBM (device pointers handed to the C ABI),
          timed with CUDA events on the stream the handle runs on; an L2 flush (256 MiB write) sits between steps,
          outside the timed intervals.
  e2e     the same metric through the public C ABI with HOST buffers (pinned): H2D of both clouds, the whole
          registration and the D2H of the pose history / statistics inside the timed region (wall clock).
  roofline  the dominant kernel's algorithmic bytes / its average launch duration measured LIVE inside one
          registration (host-stepped driver, every launch bracketed by CUDA events on the handle's stream) vs
          MEASURED_PEAKS.json's HBM copy number; isolated re-runs (ppcr_time_kernel, L2 flushed between launches) of
          the first search, a search after a cloud move, the evaluation, the tree build and the voxel filter beside it.
  cpu_baseline  the CPU oracle (a restatement of the reference; the reference itself needs PCL/Ceres which are not
          installable here) on a bounded sample of the same workload, all host cores; a 1-thread row beside it.

--impl reference times that CPU restatement instead (rank 0 only; every host core whatever OMP_NUM_THREADS says): K
bounded samples, one FULL registration and a 1-thread sample.
"""
from __future__ import annotations

import argparse
import json
import mmap
import os
import subprocess
import sys
import threading
import traceback
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "correspondences_per_sec"
UNIT = "correspondences/s"
DTYPE = "f32 rows / f64 sums"  # what the timed default path computes in; --exact runs float64 per correspondence
L2_NOTE = "256 MiB flush write between steps; working set (neighbour planes) also exceeds L2"

WORKLOADS = {
    # name: (description, params, generator kwargs)
    "c3": dict(desc="BASELINE configs[2]: 1M-pt synthetic LiDAR-like scan pair (128 rings x 7813 az), "
                    "max_neighbours=10, radius=0.5, t-dist dof=5",
               params=dict(max_neighbours=10, radius=0.5, dof=5.0)),
    "c1": dict(desc="BASELINE configs[0]: 10k-pt plane+sphere, 10deg/0.1m + noise, CLI defaults r=3 m=20 dof=5",
               params=dict(max_neighbours=20, radius=3.0, dof=5.0)),
    "c5": dict(desc="BASELINE configs[4] (one pair of): 120k-pt KITTI-like pair, CLI defaults r=3 m=20 dof=5",
               params=dict(max_neighbours=20, radius=3.0, dof=5.0)),
}
C4_PARAMS = dict(max_neighbours=10, radius=0.5, dof=5.0)  # same flags as c3
C5_POINTS = 64 * 1875                                     # points per cloud of a c5 pair


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_pair(workload: str, rank: int):
    from probabilistic_point_clouds_registration_b200 import synth
    if workload == "c3":
        src, tgt, _ = synth.config3_lidar_1m(seed=3 + rank)
    elif workload == "c1":
        src, tgt, _ = synth.config1_plane_sphere(seed=1 + rank)
    elif workload == "c5":
        src, tgt, _ = synth.config5_pair(rank)
    else:
        raise SystemExit(f"unknown workload {workload}")
    return np.ascontiguousarray(src), np.ascontiguousarray(tgt)


def static_config(workload: str, n_src: int, n_tgt: int) -> dict:
    """The `config` object: identical in both arms (what ran, nothing measured)."""
    wl = WORKLOADS[workload]
    return {"workload": wl["desc"], **wl["params"], "n_src": int(n_src), "n_tgt": int(n_tgt), "l2": L2_NOTE}


# ---------------------------------------------------------------------------------------------------------------
# synthetic inputs of the partitioned configs, generated by a fork pool BEFORE CUDA exists in this process
# ---------------------------------------------------------------------------------------------------------------

_SHARED = {}  # name -> numpy view of an anonymous shared mapping (inherited by the pool's children)


def _shared_array(name, shape):
    n_bytes = int(np.prod(shape)) * 4
    mm = mmap.mmap(-1, max(n_bytes, mmap.PAGESIZE))
    arr = np.frombuffer(mm, dtype=np.float32, count=int(np.prod(shape))).reshape(shape)
    _SHARED[name] = arr
    _SHARED[name + "/mmap"] = mm
    return arr


def _gen_c5(job):
    slot, index = job
    from probabilistic_point_clouds_registration_b200 import synth
    src, tgt, _ = synth.config5_pair(index)
    _SHARED["c5"][slot, 0] = src
    _SHARED["c5"][slot, 1] = tgt
    return slot


def _gen_c4(shape):
    from probabilistic_point_clouds_registration_b200 import synth
    rings, az = shape
    src, tgt, _ = synth.lidar_pair(4, rings, az)
    _SHARED["c4"][0] = src
    _SHARED["c4"][1] = tgt
    return -1


def generate_partitioned_inputs(args, rank, world):
    """c5: this rank's block of the batch; c4: rank 0 generates (the others receive it over NCCL later)."""
    import multiprocessing as mp
    from probabilistic_point_clouds_registration_b200 import multi
    t0 = time.perf_counter()
    mine = multi.deal_pairs(args.batch_pairs, rank, world) if args.batch_pairs > 0 else []
    extra = []
    if rank == 0 and world > 1 and mine:
        # rank 0's single-GPU reference leg runs on a share of the batch: its own block first, then the pairs that follow
        want = min(args.batch_pairs, max(len(mine), args.batch_single))
        extra = list(range(len(mine), want))
    jobs = [(slot, idx) for slot, idx in enumerate(mine + extra)]
    if jobs:
        _shared_array("c5", (len(jobs), 2, C5_POINTS, 4))
    c4_points = args.sharded_rings * args.sharded_az
    if args.sharded_rings > 0:
        _shared_array("c4", (2, c4_points, 4))
    tasks = []
    procs = max(1, min(host_threads() // max(world, 1), 32))
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=procs) as pool:
        if args.sharded_rings > 0 and rank == 0:
            tasks.append(pool.apply_async(_gen_c4, ((args.sharded_rings, args.sharded_az),)))  # the long one first
        tasks += [pool.apply_async(_gen_c5, (j,)) for j in jobs]
        for t in tasks:
            t.get()
    return {"c5_local": len(mine), "c5_total_slots": len(jobs), "gen_s": time.perf_counter() - t0, "gen_procs": procs}


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the restated reference (oracle) on a bounded sample
# ---------------------------------------------------------------------------------------------------------------

def cpu_sample(src, tgt, params: dict, n_outer: int, threads: int):
    """`n_outer` outer iterations of the reference algorithm (search + inner LM solves + cloud move) on `threads` OpenMP
    threads (an explicit count: torchrun exports OMP_NUM_THREADS=1).  Returns (correspondences, seconds, outer iterations run)."""
    from oracle import oracle as O
    O.build()
    p = O.make_params(n_iter=n_outer, **params)
    t0 = time.perf_counter()
    res = O.align(src, tgt, p, O.make_options(inner_kind=1, num_threads=threads), use_grid=True)
    dt = time.perf_counter() - t0
    corr = int(sum(s["n_correspondences"] for s in res.stats))
    return corr, dt, len(res.stats)


def cpu_rows(src, tgt, params: dict, n_outer: int, one_thread_outer: int):
    """The two CPU rows BASELINE.md section 3 promises, on bounded samples: every host core, and one thread."""
    threads = host_threads()
    corr, dt, _ = cpu_sample(src, tgt, params, n_outer, threads)
    rows = {"all": {"value": corr / dt, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"{n_outer} outer iteration(s) of the same pair (search + inner LM + cloud move) in {dt:.1f} s; "
                              f"restated reference, the real one needs PCL/Ceres"}}
    if one_thread_outer > 0:
        corr1, dt1, _ = cpu_sample(src, tgt, params, one_thread_outer, 1)
        rows["one"] = {"value": corr1 / dt1, "unit": UNIT, "cores": 1, "kind": "port",
                       "sample": f"{one_thread_outer} outer iteration(s) of the same pair on ONE thread in {dt1:.1f} s (the reference's "
                                 f"search, problem assembly and weight callback are serial)"}
    return rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    src, tgt = make_pair(args.workload, 0)
    params = WORKLOADS[args.workload]["params"]
    threads = host_threads()
    n_outer = args.cpu_outer
    for _ in range(args.warmup):
        cpu_sample(src, tgt, params, 1, threads)
    corr_total, t_total = 0, 0.0
    for _ in range(args.steps):
        corr, dt, _ = cpu_sample(src, tgt, params, n_outer, threads)
        corr_total += corr
        t_total += dt
    value = corr_total / t_total
    sample = (f"{n_outer} outer iteration(s) of the {args.workload} pair per step (search + inner LM + cloud move) on {threads} "
              f"threads; the full registration and a 1-thread sample are reported beside it")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(args.workload, len(src), len(tgt)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_ref_full:
        # one FULL registration (constructor + align() to the reference's stopping rule) on every host core
        corr, dt, outer = cpu_sample(src, tgt, params, 1000, threads)
        line["full_registration"] = {"seconds": dt, "outer_iterations": outer, "correspondences": corr, "value": corr / dt,
                                     "unit": UNIT, "cores": threads}
        corr1, dt1, _ = cpu_sample(src, tgt, params, 1, 1)
        line["cpu_baseline_1thread"] = {"value": corr1 / dt1, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": f"1 outer iteration of the same pair on ONE thread in {dt1:.1f} s"}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------

def _pin(arr, chunk_bytes):
    """cudaHostRegister of a numpy array's memory (the shared mappings the inputs were generated into), in chunks: one
    registration of the whole 3.9 GB batch is refused on some boxes (cudaErrorOperatingSystem), and a refused call must not
    leave its error behind for the next CUDA call of this thread.  A copy must not straddle two registrations (CUDA refuses
    it), so chunk_bytes has to be a multiple of whatever is copied in one call.  Returns the fraction of the bytes pinned."""
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so.12")
    except OSError:
        return 0.0
    rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
    base, total, done = arr.ctypes.data, arr.nbytes, 0
    while done < total:
        n = min(chunk_bytes, total - done)
        if rt.cudaHostRegister(base + done, n, 0) != 0:
            rt.cudaGetLastError()  # clear it
            break
        done += n
    return done / max(total, 1)


def bench_batch_c5(args, ctx):
    """BASELINE configs[4]: the 1024-pair batch, strong scaling.  Returns the sub-record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from probabilistic_point_clouds_registration_b200 import capi
    rank, world, local_rank = ctx["rank"], ctx["world"], ctx["local_rank"]
    buf = _SHARED.get("c5")
    n_local = ctx["gen"]["c5_local"]
    if buf is None or args.batch_pairs <= 0:
        return None
    params = capi.make_params(**WORKLOADS["c5"]["params"])
    el = C5_POINTS * 4 * 4  # bytes per cloud
    d_buf = torch.from_numpy(buf).cuda()  # resident copy of every slot this rank may run (before pinning: one copy of it all)
    torch.cuda.synchronize()
    pinned = _pin(buf, 64 * el)  # 32 pairs per registration; ppcr copies cloud by cloud
    torch.cuda.synchronize()

    def host_pairs(slots):
        return [(buf[s, 0], buf[s, 1]) for s in slots]

    def dev_pairs(slots):
        base = d_buf.data_ptr()
        return [(base + (2 * s) * el, C5_POINTS, base + (2 * s + 1) * el, C5_POINTS) for s in slots]

    opt_host = capi.make_options(device=local_rank)
    opt_dev = capi.make_options(device=local_rank, input_on_device=True)

    def timed(pairs, opt):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx["barrier"]()
        t0 = time.perf_counter()
        ev0.record()
        T, n_outer, corr = capi.align_batch(pairs, params, opt, slots=args.batch_slots) if pairs else (np.zeros((0, 4, 4)), np.zeros(0), np.zeros(0))
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        return T, n_outer, corr, ev0.elapsed_time(ev1) * 1e-3, wall

    mine = list(range(n_local))
    warm = mine[:min(len(mine), 2 * max(args.batch_slots, 6))]
    if warm:
        capi.align_batch(dev_pairs(warm), params, opt_dev, slots=args.batch_slots)
    # every leg `batch_reps` times (max over ranks per repetition, the median repetition is reported, all are listed): on the
    # shared boxes one run of a batch in ten or so takes up to twice as long, whatever is being measured
    dev_runs, host_runs = [], []
    for _ in range(max(1, args.batch_reps)):
        T_dev, outer_dev, corr_dev, s, _ = timed(dev_pairs(mine), opt_dev)
        dev_runs.append(s)
        T_host, outer_host, corr_host, s, wall_host = timed(host_pairs(mine), opt_host)
        host_runs.append(s)
    same = bool(np.array_equal(T_dev, T_host))  # the same pairs from host or device buffers: bit-identical poses
    runs = torch.tensor([dev_runs, host_runs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(runs, op=dist.ReduceOp.MAX)
    dev_runs, host_runs = ([float(v) for v in row] for row in runs.tolist())
    s_dev, s_host = float(np.median(dev_runs)), float(np.median(host_runs))
    c = torch.tensor([int(corr_dev.sum()), int(outer_dev.sum()), int(same), len(mine)], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    corr_all, outer_all, same_all, n_all = (int(v) for v in c.tolist())
    single = None
    if world > 1:
        # rank 0 alone on a share of the batch, every other GPU idle: the single-GPU rate in the same process and box
        ctx["barrier"]()
        if rank == 0:
            slots = list(range(ctx["gen"]["c5_total_slots"]))
            secs = []
            for _ in range(max(1, args.batch_reps)):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                capi.align_batch(host_pairs(slots), params, opt_host, slots=args.batch_slots)
                ev1.record()
                torch.cuda.synchronize()
                secs.append(ev0.elapsed_time(ev1) * 1e-3)
            single = {"pairs": len(slots), "e2e_pairs_per_s": len(slots) / float(np.median(secs)),
                      "e2e_seconds_all": [round(v, 4) for v in secs]}
        ctx["barrier"]()
    if rank != 0:
        return None
    rec = {
        "workload": f"BASELINE configs[4]: {n_all} independent 120k-pt pairs (seeds 1000..), CLI defaults m=20 r=3 dof=5, "
                    f"dealt in contiguous blocks to {world} rank(s), ppcr_align_batch with {args.batch_slots} lanes per rank",
        "n_pairs": n_all, "n_gpus": world, "scaling": "strong",
        "pairs_per_s": n_all / s_dev, "e2e_pairs_per_s": n_all / s_host,
        "seconds": s_dev, "e2e_seconds": s_host, "seconds_all": [round(v, 4) for v in dev_runs],
        "e2e_seconds_all": [round(v, 4) for v in host_runs],
        "correspondences_per_s": corr_all / s_dev, "mean_outer_iterations": outer_all / max(n_all, 1),
        "h2d_bytes_per_pair": 2 * el, "d2h_bytes_per_pair": 16 * 8 + 4 + 8, "host_buffers_pinned_fraction": pinned,
        "host_vs_device_inputs_bit_identical": same_all == world,
        "timing": "CUDA events on the rank's current stream around ppcr_align_batch (it returns when every lane has finished), "
                  "max over ranks",
    }
    if single:
        rec["single_gpu"] = single
        rec["speedup_vs_single_gpu"] = rec["e2e_pairs_per_s"] / single["e2e_pairs_per_s"]
    return rec


def bench_sharded_c4(args, ctx):
    """BASELINE configs[3]: one 10M-point pair sharded over the ranks, with the on-box parity assertion."""
    import torch
    import torch.distributed as dist
    from probabilistic_point_clouds_registration_b200 import capi, multi
    rank, world, local_rank = ctx["rank"], ctx["world"], ctx["local_rank"]
    buf = _SHARED.get("c4")
    if buf is None:
        return None
    n = buf.shape[1]
    d_all = torch.empty((2, n, 4), dtype=torch.float32, device="cuda")
    if rank == 0:
        d_all.copy_(torch.from_numpy(buf))
    if world > 1:
        dist.broadcast(d_all, src=0)  # set-up only: the data path of the registration has no collective
    torch.cuda.synchronize()
    d_src, d_tgt = d_all[0], d_all[1]
    params = capi.make_params(**C4_PARAMS)
    opt = capi.make_options(device=local_rank, input_on_device=True)

    def one_gpu():
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        with capi.Registration(d_src.data_ptr(), d_tgt.data_ptr(), params, opt, n_source=n, n_target=n) as reg:
            reg.align()
            ev1.record()
            hist, stats = reg.transformation_history(), reg.iteration_stats()
        torch.cuda.synchronize()
        return hist, stats, ev0.elapsed_time(ev1)

    one_gpu()  # warm-up (pool growth, graph instantiation)
    single_ms = []
    for _ in range(args.sharded_reps):
        ref_hist, ref_stats, ms = one_gpu()
        single_ms.append(ms)
    rec = {"workload": f"BASELINE configs[3]: one {n}-pt pair (seed 4), -m 10 -r 0.5 -d 5, source dealt block-cyclically ({args.sharded_block}-point runs) to {world} rank(s), "
                       f"target octree replicated, 25-double moment exchange written peer to peer from inside k_evalctl",
           "n_gpus": world, "n_src": n, "n_tgt": n, "outer_iterations": len(ref_stats),
           "correspondences": int(sum(s["n_correspondences"] for s in ref_stats)),
           "single_gpu_ms": float(np.median(single_ms)), "single_gpu_ms_all": [round(v, 2) for v in single_ms]}
    if world == 1:
        rec["ms_per_registration"] = rec["single_gpu_ms"]
        rec["sharded_parity"] = "n/a (one GPU)"
        return rec
    # the source is dealt block-cyclically (runs of 8192 consecutive points, round robin): every rank gets arcs of every ring,
    # so the search load is even (contiguous slices: 189 against 151 ms of search on two GPUs, the near rings being denser)
    mine_idx = torch.from_numpy(multi.block_cyclic_indices(n, rank, world, args.sharded_block)).cuda()
    d_mine = d_src.index_select(0, mine_idx).contiguous()
    torch.cuda.synchronize()
    src_ptr, n_mine = d_mine.data_ptr(), int(d_mine.shape[0])
    times, exchange, phases, parity = [], [], [], "ok"
    import gc
    gc.collect()
    gc.disable()  # (a generation-2 collection inside the read-back phase would be timed as part of a registration)
    for rep in range(args.sharded_reps + 1):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx["barrier"]()
        ev0.record()
        w0 = time.perf_counter()
        with multi.ShardedRegistration(src_ptr, d_tgt.data_ptr(), params, rank, world, opt, n_source=n_mine, n_target=n) as reg:
            w1 = time.perf_counter()
            reg.align()  # (returns after its own read-back of the final state: pose, iteration count, error flag)
            ev1.record()
            w2 = time.perf_counter()
            hist, stats = reg.transformation_history(), reg.iteration_stats()
            lt = reg.stage_times()
            w2b = time.perf_counter()
        w3 = time.perf_counter()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1), lt.exchange_wait_ms, 1e3 * (w1 - w0), 1e3 * (w2 - w1), 1e3 * (w2b - w2), 1e3 * (w3 - w2b)],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep > 0:
            times.append(float(t[0].item()))
            exchange.append((float(t[1].item()), int(lt.exchanges)))
            phases.append([round(float(v), 2) for v in t[2:].tolist()])
        # parity, asserted on the box: bit-identical histories on every rank, and the single-GPU pose
        mine = torch.from_numpy(np.ascontiguousarray(hist)).cuda()
        shapes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(shapes, torch.tensor([mine.shape[0]], dtype=torch.int64, device="cuda"))
        if len({int(s.item()) for s in shapes}) != 1:
            parity = f"FAILED: ranks ran different numbers of outer iterations {[int(s.item()) for s in shapes]}"
        else:
            got = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(got, mine)
            if not all(bool((g == mine).all()) for g in got):
                parity = "FAILED: pose histories differ between ranks"
        if parity == "ok":
            if len(hist) != len(ref_hist):
                parity = f"FAILED: {len(hist)} outer iterations sharded vs {len(ref_hist)} on one GPU"
            else:
                dR = hist[-1][:3, :3].T @ ref_hist[-1][:3, :3]
                ang = float(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
                dt = float(np.linalg.norm(hist[-1][:3, 3] - ref_hist[-1][:3, 3]))
                k_s = [s["n_correspondences"] for s in stats]
                k_r = [s["n_correspondences"] for s in ref_stats]
                rec["pose_delta_vs_single_gpu"] = {"rad": ang, "m": dt, "association_sizes_equal": k_s == k_r}
                if not (ang < 1e-6 and dt < 1e-6):
                    parity = f"FAILED: pose differs from the single-GPU run by {ang:.2e} rad / {dt:.2e} m"
    gc.enable()
    rec["ms_per_registration"] = float(np.median(times))
    rec["ms_all"] = [round(v, 2) for v in times]
    rec["phases_ms_all"] = {"constructor_incl_token_exchange, align, read_back, destroy (host clock, max over ranks)": phases}
    rec["speedup_vs_single_gpu"] = rec["single_gpu_ms"] / rec["ms_per_registration"]
    rec["sharded_parity"] = parity
    if exchange:
        worst = max(e[0] for e in exchange)
        rec["exchange"] = {"per_registration_ms_max_over_ranks": worst, "exchanges": exchange[-1][1],
                           "us_per_exchange": 1e3 * worst / max(exchange[-1][1], 1),
                           "note": "SM clocks between the controller block starting to send its 25 doubles and having every peer's "
                                   "(includes waiting for the slowest rank's evaluation to finish: load imbalance shows up here)"}
    rec["timing"] = ("CUDA events around constructor (incl. the token exchange) + align() (which ends with its own read-back of the final "
                     "state), clouds resident, max over ranks; the history / statistics reads that follow are outside (phases_ms_all lists them)")
    return rec if rank == 0 else None


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # synthetic inputs first: a fork pool must not inherit a CUDA context
    wl = WORKLOADS[args.workload]
    partitioned = args.workload == "c3" and not args.headline_only
    gen = {"c5_local": 0, "c5_total_slots": 0, "gen_s": 0.0, "gen_procs": 0}
    if partitioned:
        gen = generate_partitioned_inputs(args, rank, world)
    src, tgt = make_pair(args.workload, rank)
    n_src, n_tgt = len(src), len(tgt)

    import torch
    import torch.distributed as dist

    from probabilistic_point_clouds_registration_b200 import build, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the registration path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    build.build_cuda()
    capi.lib()

    params = capi.make_params(**wl["params"])
    stream = torch.cuda.Stream()
    # device-resident copies (for `value`) and pinned host copies (for `e2e`)
    d_src = torch.from_numpy(src).cuda()
    d_tgt = torch.from_numpy(tgt).cuda()
    h_src = torch.from_numpy(src).pin_memory()
    h_tgt = torch.from_numpy(tgt).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = {"rank": rank, "world": world, "local_rank": local_rank, "barrier": barrier, "gen": gen}

    def register_device(record_events=None, exact=False):
        """ctor + align() + read-back of the result with the clouds already in HBM."""
        opt = capi.make_options(device=local_rank, input_on_device=True, stream=stream.cuda_stream, exact_weights=exact)
        if record_events:
            record_events[0].record(stream)
        reg = capi.Registration(d_src.data_ptr(), d_tgt.data_ptr(), params, opt, n_source=n_src, n_target=n_tgt)
        reg.align()
        if record_events:
            record_events[1].record(stream)
        return reg

    def register_host():
        opt = capi.make_options(device=local_rank, stream=stream.cuda_stream)
        reg = capi.Registration(h_src.numpy(), h_tgt.numpy(), params, opt)
        reg.align()
        hist = reg.transformation_history()
        stats = reg.iteration_stats()
        return reg, hist, stats

    with torch.cuda.stream(stream):
        # ---- warm-up ------------------------------------------------------------------------------------------
        for _ in range(max(args.warmup, 3)):
            reg = register_device()
            reg.close()
        # ---- `value`: K timed steps, device-resident inputs -----------------------------------------------------
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        corr_total, launches, n_outer_last, stats_last = 0, 0, 0, None
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            flush.fill_(k & 0xff)  # evict L2 between steps (outside the timed interval)
            reg = register_device(ev[k])
            stats_last = reg.iteration_stats()
            corr_total += sum(s["n_correspondences"] for s in stats_last)
            launches += reg.stage_times().total_launches
            n_outer_last = len(stats_last)
            if k == args.steps - 1:
                keep = reg
            else:
                reg.close()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # ---- `e2e`: K timed steps through the C ABI with host buffers -------------------------------------------
        for _ in range(2):
            r, _, _ = register_host()
            r.close()
        barrier()
        e2e_corr = 0
        d2h_bytes = 0
        e2e_steps_ms = []
        t0 = time.perf_counter()
        for k in range(args.steps):
            tk = time.perf_counter()
            r, hist, stats = register_host()
            e2e_corr += sum(s["n_correspondences"] for s in stats)
            d2h_bytes = hist.nbytes + 40 * len(stats)
            r.close()
            e2e_steps_ms.append(1e3 * (time.perf_counter() - tk))
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop()
        # ---- the float64-per-correspondence mode (options.exact_weights) beside the default, same pair ----------
        exact_ms = None
        if rank == 0 and args.exact_steps > 0:
            register_device(exact=True).close()
            evx = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.exact_steps)]
            for k in range(args.exact_steps):
                register_device(evx[k], exact=True).close()
            torch.cuda.synchronize()
            exact_ms = sum(a.elapsed_time(b) for a, b in evx) / args.exact_steps
        # ---- roofline of the dominant kernels, in isolation, on the last handle's final state -------------------
        # k_search runs once from scratch and (outer - 1) times fused with the cloud move and warm-started; both forms
        # are timed, the roofline entry is their launch-weighted mean
        kernels = {}
        for which, name in ((0, "k_search_first"), (4, "k_search_moved"), (1, "k_evalctl"), (3, "tree_build")):
            ms, nbytes = keep.time_kernel(which, reps=10, flush_l2=True)
            kernels[name] = {"avg_ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9}
        # ---- the same kernels timed LIVE inside one registration: the host-stepped driver brackets every launch with
        # CUDA events on the handle's stream (the device-side WHILE graph of the product path cannot be bracketed)
        opt = capi.make_options(device=local_rank, input_on_device=True, stream=stream.cuda_stream, driver=1,
                                record_stage_times=True)
        live = capi.Registration(d_src.data_ptr(), d_tgt.data_ptr(), params, opt, n_source=n_src, n_target=n_tgt)
        live.align()
        lt = live.stage_times()
        live_stats = live.iteration_stats()
        live.close()
        in_loop = {
            # the host-stepped driver launches k_search every tick; all but one per outer iteration exit at once (a few
            # microseconds each, left in the sum: the figure errs on the slow side)
            "k_search": {"avg_ms": lt.search_ms / max(len(live_stats), 1), "launches": len(live_stats)},
            "k_evalctl": {"avg_ms": lt.eval_ms / max(lt.eval_launches, 1), "launches": lt.eval_launches},
        }
        evals = sum(s["lm_iterations"] + 1 for s in stats_last)
        n_moved = max(n_outer_last - 1, 0)
        launches_of = {"k_search_first": min(n_outer_last, 1), "k_search_moved": n_moved, "k_evalctl": evals, "tree_build": 1}
        share = {k: kernels[k]["avg_ms"] * launches_of[k] for k in kernels}
        n_s = max(n_outer_last, 1)
        kernels["k_search"] = {
            "avg_ms": (share["k_search_first"] + share["k_search_moved"]) / n_s,
            "algorithmic_bytes": (kernels["k_search_first"]["algorithmic_bytes"] * launches_of["k_search_first"]
                                  + kernels["k_search_moved"]["algorithmic_bytes"] * n_moved) / n_s}
        kernels["k_search"]["gbs"] = kernels["k_search"]["algorithmic_bytes"] / (kernels["k_search"]["avg_ms"] * 1e-3) / 1e9
        share = {"k_search": share["k_search_first"] + share["k_search_moved"], "k_evalctl": share["k_evalctl"],
                 "tree_build": share["tree_build"]}
        launches_of["k_search"] = n_s
        # in-loop figures replace the isolated ones where the live pass has them (same algorithmic bytes per launch)
        for name in ("k_search", "k_evalctl"):
            if in_loop[name]["launches"] > 0 and in_loop[name]["avg_ms"] > 0:
                kernels[name]["isolated_avg_ms"] = kernels[name]["avg_ms"]
                kernels[name]["avg_ms"] = in_loop[name]["avg_ms"]
                kernels[name]["gbs"] = kernels[name]["algorithmic_bytes"] / (kernels[name]["avg_ms"] * 1e-3) / 1e9
                share[name] = in_loop[name]["avg_ms"] * in_loop[name]["launches"]
        keep.close()
        # voxel filter (not on config 3's path: BASELINE configs[1] filters both clouds at 0.05 m): the source of this pair
        # through ppcr_voxel_filter at that leaf, timed end to end on the device
        if rank == 0 and not args.headline_only:
            vox = capi.voxel_filter_timed(d_src.data_ptr(), n_src, 0.05, device=local_rank, reps=3)
            if vox:
                kernels["voxel_filter"] = vox

    # max over ranks of the timed durations; sums over ranks of the work
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, t_wall = (float(v) for v in t.tolist())
        c = torch.tensor([corr_total, e2e_corr, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        corr_total, e2e_corr, launches = (int(v) for v in c.tolist())

    # ---- the partitioned configs at this N ----------------------------------------------------------------------
    batch_rec = sharded_rec = None
    if partitioned:
        del flush
        torch.cuda.empty_cache()
        try:
            batch_rec = bench_batch_c5(args, ctx)
        except Exception as e:  # the headline line must still print
            traceback.print_exc()
            batch_rec = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None
        try:
            sharded_rec = bench_sharded_c4(args, ctx)
        except Exception as e:
            traceback.print_exc()
            sharded_rec = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        dom = max(("k_search", "k_evalctl"), key=lambda k: share[k])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                    "avg_launch_ms": kernels[dom]["avg_ms"],
                    "kernels": {k: {**v, "frac": v["gbs"] / peak, "launches_per_step": launches_of.get(k)}
                                for k, v in kernels.items()},
                    "share_of_step_ms": share,
                    "note": "k_search / k_evalctl: average launch duration inside one live registration (host-stepped "
                            "driver, every launch bracketed by CUDA events on the handle's stream); *_first / *_moved, "
                            "tree_build, voxel_filter and isolated_avg_ms: isolated re-runs with a 256 MiB L2 flush before "
                            "each launch.  The search is an octree walk (issue bound), not a stream: see DESIGN.md 4.1 "
                            "and profiles/"}
        cpu, cpu1 = None, None
        if not args.no_cpu:
            rows = cpu_rows(src, tgt, wl["params"], args.cpu_outer, 0 if args.headline_only else 1)
            cpu, cpu1 = rows["all"], rows.get("one")
        line = {
            "metric": METRIC, "value": corr_total / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": static_config(args.workload, n_src, n_tgt),
            "run": {"pairs_per_step": world, "outer_iterations": n_outer_last,
                    "correspondences_per_pair": corr_total // max(1, args.steps * world),
                    "exact_weights_ms_per_step": exact_ms, "input_generation_s": gen["gen_s"]},
            "clocks": clocks,
            "e2e": {"value": e2e_corr / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "ms_per_step_median": float(np.median(e2e_steps_ms)), "ms_all": [round(v, 2) for v in e2e_steps_ms],
                    "h2d_bytes_per_step": int(src.nbytes + tgt.nbytes), "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "cpu_baseline_1thread": cpu1,
            "batch_c5": batch_rec,
            "sharded_c4": sharded_rec,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-outer", type=int, default=2, help="outer iterations in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref-full", action="store_true", help="--impl reference: skip the full registration and the 1-thread row")
    ap.add_argument("--headline-only", action="store_true", help="skip the batch_c5 / sharded_c4 sub-records and the extra rows")
    ap.add_argument("--exact-steps", type=int, default=3, help="steps of the exact_weights=1 timing beside the default")
    ap.add_argument("--batch-pairs", type=int, default=1024, help="pairs of the c5 batch (BASELINE configs[4])")
    ap.add_argument("--batch-slots", type=int, default=6, help="lanes per rank of ppcr_align_batch")
    ap.add_argument("--batch-reps", type=int, default=3, help="repetitions of each leg of the batch (the median is reported)")
    ap.add_argument("--batch-single", type=int, default=192, help="N > 1: pairs of rank 0's single-GPU reference leg")
    ap.add_argument("--sharded-rings", type=int, default=320, help="rings of the sharded pair (320 x 31250 = 10M points); 0 = skip")
    ap.add_argument("--sharded-az", type=int, default=31250)
    ap.add_argument("--sharded-reps", type=int, default=5)
    ap.add_argument("--sharded-block", type=int, default=8192, help="points per run of the block-cyclic deal of the sharded pair")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), *sys.argv[1:]]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
