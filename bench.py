#!/usr/bin/env python
"""bench.py -- the registration hot path on 1..N B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c5] [--impl ours|reference]

A "step" is one registration of one synthetic scan pair: ProbPointCloudRegistration constructor + align() to the
reference's own stopping rule (src/prob_point_cloud_registration.cc:15-158) -- octree build, then per outer
iteration radius search -> weights + normal equations per LM iteration -> pose update -> cloud move -> convergence
test, all on the device.  The default workload is BASELINE.json configs[2], the 1M-point pair the metric is quoted
on ("c3": -m 10 -r 0.5 -d 5).  With N > 1 every rank registers its own pair of the same shape (independent scan
pairs, data-parallel, no data-path collective): weak scaling.

  value   correspondences/s with the clouds already resident in HBM (device pointers handed to the C ABI),
          timed with CUDA events on the stream the handle runs on; an L2 flush (256 MiB write) sits between steps,
          outside the timed intervals.
  e2e     the same metric through the public C ABI with HOST buffers (pinned): H2D of both clouds, the whole
          registration and the D2H of the pose history / statistics inside the timed region (wall clock).
  roofline  the dominant kernel's algorithmic bytes / its average launch duration measured LIVE inside one
          registration (host-stepped driver, every launch bracketed by CUDA events on the handle's stream) vs
          MEASURED_PEAKS.json's HBM copy number; isolated re-runs (ppcr_time_kernel, L2 flushed between launches) of
          the first search, a search after a cloud move and the evaluation are reported beside it.
  cpu_baseline  the CPU oracle (a restatement of the reference; the reference itself needs PCL/Ceres which are not
          installable here) on a bounded sample of the same workload, all host cores.

--impl reference times that CPU restatement instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "correspondences_per_sec"
UNIT = "correspondences/s"

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

def cpu_sample(src, tgt, params: dict, n_outer: int):
    """`n_outer` outer iterations of the reference algorithm (search + inner LM solves + cloud move) with all host
    threads.  Returns (correspondences, seconds, threads)."""
    from oracle import oracle as O
    O.build()
    p = O.make_params(n_iter=n_outer, **params)
    t0 = time.perf_counter()
    res = O.align(src, tgt, p, O.make_options(inner_kind=1, num_threads=0), use_grid=True)
    dt = time.perf_counter() - t0
    corr = int(sum(s["n_correspondences"] for s in res.stats))
    return corr, dt, O.max_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    src, tgt = make_pair(args.workload, 0)
    params = WORKLOADS[args.workload]["params"]
    n_outer = args.cpu_outer
    for _ in range(args.warmup):
        cpu_sample(src, tgt, params, 1)
    corr_total, t_total, threads = 0, 0.0, 1
    for _ in range(args.steps):
        corr, dt, threads = cpu_sample(src, tgt, params, n_outer)
        corr_total += corr
        t_total += dt
    value = corr_total / t_total
    sample = (f"{n_outer} outer iteration(s) of the {args.workload} pair per step (search + inner LM + cloud move), "
              f"not the full registration")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["desc"], **params, "n_src": len(src), "n_tgt": len(tgt)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    from probabilistic_point_clouds_registration_b200 import build, capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the registration path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    build.build_cuda()
    capi.lib()

    wl = WORKLOADS[args.workload]
    params = capi.make_params(**wl["params"])
    src, tgt = make_pair(args.workload, rank)
    n_src, n_tgt = len(src), len(tgt)

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

    def register_device(record_events=None):
        """ctor + align() + read-back of the result with the clouds already in HBM."""
        opt = capi.make_options(device=local_rank, input_on_device=True, stream=stream.cuda_stream)
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
        t0 = time.perf_counter()
        for k in range(args.steps):
            tk = time.perf_counter()
            r, hist, stats = register_host()
            e2e_corr += sum(s["n_correspondences"] for s in stats)
            d2h_bytes = hist.nbytes + 40 * len(stats)
            r.close()
            if os.environ.get("PPCR_BENCH_DEBUG"):
                print(f"[rank {rank}] e2e step {k}: {1e3 * (time.perf_counter() - tk):.2f} ms", file=sys.stderr)
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop()
        # ---- roofline of the dominant kernels, in isolation, on the last handle's final state -------------------
        # k_search runs once from scratch and (outer - 1) times fused with the cloud move and warm-started; both forms
        # are timed, the roofline entry is their launch-weighted mean
        kernels = {}
        for which, name in ((0, "k_search_first"), (4, "k_search_moved"), (1, "k_evalctl")):
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
        launches_of = {"k_search_first": min(n_outer_last, 1), "k_search_moved": n_moved, "k_evalctl": evals}
        share = {k: kernels[k]["avg_ms"] * launches_of[k] for k in kernels}
        n_s = max(n_outer_last, 1)
        kernels["k_search"] = {
            "avg_ms": (share["k_search_first"] + share["k_search_moved"]) / n_s,
            "algorithmic_bytes": (kernels["k_search_first"]["algorithmic_bytes"] * launches_of["k_search_first"]
                                  + kernels["k_search_moved"]["algorithmic_bytes"] * n_moved) / n_s}
        kernels["k_search"]["gbs"] = kernels["k_search"]["algorithmic_bytes"] / (kernels["k_search"]["avg_ms"] * 1e-3) / 1e9
        share = {"k_search": share["k_search_first"] + share["k_search_moved"], "k_evalctl": share["k_evalctl"]}
        launches_of["k_search"] = n_s
        # in-loop figures replace the isolated ones where the live pass has them (same algorithmic bytes per launch)
        for name in ("k_search", "k_evalctl"):
            if in_loop[name]["launches"] > 0 and in_loop[name]["avg_ms"] > 0:
                kernels[name]["isolated_avg_ms"] = kernels[name]["avg_ms"]
                kernels[name]["avg_ms"] = in_loop[name]["avg_ms"]
                kernels[name]["gbs"] = kernels[name]["algorithmic_bytes"] / (kernels[name]["avg_ms"] * 1e-3) / 1e9
                share[name] = in_loop[name]["avg_ms"] * in_loop[name]["launches"]
        keep.close()

    # max over ranks of the timed durations; sums over ranks of the work
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, t_wall = (float(v) for v in t.tolist())
        c = torch.tensor([corr_total, e2e_corr, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        corr_total, e2e_corr, launches = (int(v) for v in c.tolist())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        dom = max(share, key=share.get)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                    "avg_launch_ms": kernels[dom]["avg_ms"],
                    "kernels": {k: {**v, "frac": v["gbs"] / peak, "launches_per_step": launches_of[k]}
                                for k, v in kernels.items()},
                    "share_of_step_ms": share,
                    "note": "k_search / k_evalctl: average launch duration inside one live registration (host-stepped "
                            "driver, every launch bracketed by CUDA events on the handle's stream); *_first / *_moved and "
                            "isolated_avg_ms: isolated re-runs with a 256 MiB L2 flush before each launch.  The search "
                            "is an octree walk (issue bound), not a stream: see DESIGN.md 4.1 and profiles/"}
        cpu = None
        if not args.no_cpu:
            corr, dt, threads = cpu_sample(src, tgt, wl["params"], args.cpu_outer)
            cpu = {"value": corr / dt, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_outer} outer iteration(s) of the same pair (search + inner LM + cloud move) "
                             f"in {dt:.1f} s; restated reference, the real one needs PCL/Ceres"}
        line = {
            "metric": METRIC, "value": corr_total / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], **wl["params"], "n_src": n_src, "n_tgt": n_tgt,
                       "pairs_per_step": world, "outer_iterations": n_outer_last,
                       "correspondences_per_pair": corr_total // max(1, args.steps * world),
                       "l2": "256 MiB flush write between steps; working set (neighbour planes) also exceeds L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_corr / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "h2d_bytes_per_step": int(src.nbytes + tgt.nbytes), "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
        }
        print(json.dumps(line))
    if world > 1:
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
