#!/usr/bin/env python
"""bench.py -- collision-checked configurations / second on B200 for the batched configuration-feasibility path.

  python bench.py [--gpus N --steps K --warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4] [--configs M]

A "step" is one pass of the hot path (FK -> limits -> environment + self collision) over one batch of M
synthetic configurations of the workload.  At N=1 the workload is BASELINE.json configs[1]: the 6-DOF arm in
the cluttered mesh world (200 obstacles, ~500k triangles), 1M uniform random configurations per step.  Under
torchrun (N>1) every rank holds a replica of the static geometry and checks its own M configurations (weak
scaling, no data-path collective: only timings are reduced).

  value     configs/s, inputs already resident in HBM when the timed region starts (kb_feasible_batch_device)
  e2e       the same metric through the C ABI host entry point (kb_feasible_batch) with pinned host buffers:
            H2D of the configurations and D2H of the feasibility bytes are inside the timed region
  roofline  traversal kernel: algorithmic bytes per launch / measured launch duration vs the measured HBM peak
  cpu_baseline  the CPU oracle (restatement of the reference path) on this box's host cores, bounded sample

--impl reference times the CPU path alone (oracle, all host threads; the real reference cannot be built here
because its arithmetic lives in the absent KrisLibrary) on the same workload, each step a bounded sample.
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
# all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would silently serialise the CPU arm)
NTHREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--configs", type=int, default=0, help="configurations (or edges) per step per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    return ap.parse_args()


WORKLOADS = {
    "c1": dict(name="C1 arm6 + ground + 10 boxes (self+env), uniform random configurations", n=10000),
    "c2": dict(name="C2 arm6 (TX90-class, 7 links) in cluttered mesh world: 200 blob obstacles, ~500k triangles, self+env, "
                    "uniform random configurations", n=1_000_000),
    "c3": dict(name="C3 dual-arm 15-DOF (18 links) self-collision only, uniform random configurations", n=1_000_000),
    "c4": dict(name="C4 straight-line edges in the C2 world at eps=0.01 (EpsilonEdgeChecker)", n=100_000),
    "c5": dict(name="C5 arm6 meshes vs 5M-point cloud (points on the C2 obstacle surfaces + 5 mm noise, margin 5 mm): collide bit + "
                    "min distance with upperBound 0.5 m per configuration", n=200_000),
}


def make_world(which):
    from klampt_b200 import synth
    if which == "c1":
        return synth.world_c1()
    if which in ("c2", "c4"):
        return synth.world_c2()
    if which == "c5":
        return synth.world_c5()
    return synth.world_c3()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def count(self, t0, t1=None):
        """samples read inside the window [t0, t1] (host clock at the time the line arrived)"""
        t1 = time.perf_counter() if t1 is None else t1
        return sum(1 for t, r in self.rows if t0 <= t <= t1 and len(r) >= 8)

    def stop(self, window=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if len(r) >= 8 and (window is None or window[0] <= t <= window[1])]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if r[4 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_leg(orc, gen_batch, budget_s, edges=False, slice_n=100_000, with_dist=False):
    """oracle with all host threads on successive slices of the workload until the time budget is spent"""
    from oracle.oracle import max_threads
    done, t_used, feas = 0, 0.0, 0
    k = 0
    while t_used < budget_s and k < 64:
        data = gen_batch(k, slice_n)
        t0 = time.perf_counter()
        if edges:
            out, _ = orc.edges_visible_batch(data[0], data[1], eps=0.01, nthreads=NTHREADS)
        else:
            out = orc.feasible_batch(data, nthreads=NTHREADS)
            if with_dist:
                orc.distance_batch(data, upper_bound=0.5, include_self=False, nthreads=NTHREADS)
        t_used += time.perf_counter() - t0
        done += len(out)
        feas += int(out.sum())
        k += 1
    return done / t_used, NTHREADS, done, t_used, feas / max(1, done)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    M = args.configs or wl["n"]
    edges = args.workload == "c4"
    metric = "edge checks/sec" if edges else "collision-checked configs/sec"
    unit = "edges/s" if edges else "configs/s"

    from klampt_b200 import synth
    spec = make_world(args.workload)
    robot = spec.robot

    def gen_configs(k, n):
        return synth.sample_configs(robot, n, 1000 + 17 * k + rank * 101)

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle.oracle import OracleWorld
        orc = OracleWorld(spec)
        per_step = min(M, 2_000 if edges else (10_000 if args.workload == "c5" else 200_000))
        if edges:
            A0, B0 = synth.sample_edges(robot, lambda Q: orc.feasible_batch(Q), per_step, 4)
            gen = lambda k, n: (A0, B0)
        else:
            gen = gen_configs
        times = []
        for s in range(args.warmup + args.steps):
            data = gen(s, per_step)
            t0 = time.perf_counter()
            if edges:
                orc.edges_visible_batch(data[0], data[1], eps=0.01, nthreads=NTHREADS)
            else:
                orc.feasible_batch(data, nthreads=NTHREADS)
                if args.workload == "c5":
                    orc.distance_batch(data, upper_bound=0.5, include_self=False, nthreads=NTHREADS)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
        from oracle.oracle import max_threads
        v = per_step * len(times) / sum(times)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": wl["name"], "per_step": per_step},
                "cpu_baseline": {"value": v, "unit": unit, "cores": NTHREADS, "kind": "port",
                                 "sample": "%d %s per step (bounded sample of the %d-per-step workload), oracle with OpenMP over all host threads"
                                           % (per_step, "edges" if edges else "configurations", M)},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from klampt_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(spec, device=local_rank)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    L = robot.L
    NB = 4                                    # rotating input batches: NB x M x L x 8 B  (> L2 for the default M)
    orc = None
    if edges:
        from oracle.oracle import OracleWorld
        orc = OracleWorld(spec)
        hostA, hostB = [], []
        for k in range(NB):
            A, B = synth.sample_edges(robot, lambda Q: eng.feasible_batch(Q), M, 4 + k + 10 * rank)
            hostA.append(torch.from_numpy(A).pin_memory())
            hostB.append(torch.from_numpy(B).pin_memory())
        devA = [a.cuda(non_blocking=True) for a in hostA]
        devB = [b.cuda(non_blocking=True) for b in hostB]
        host_out = torch.empty(M, dtype=torch.uint8).pin_memory()
    else:
        hostQ = [torch.from_numpy(gen_configs(k, M)).pin_memory() for k in range(NB)]
        devQ = [q.cuda(non_blocking=True) for q in hostQ]
        host_out = torch.empty(M, dtype=torch.uint8).pin_memory()
    dev_out = torch.empty(M, dtype=torch.uint8, device="cuda")
    with_dist = args.workload == "c5"
    dev_dist = torch.empty(M, dtype=torch.float64, device="cuda") if with_dist else None
    host_dist = torch.empty(M, dtype=torch.float64).pin_memory() if with_dist else None
    torch.cuda.synchronize()

    def step_device(k):
        if edges:
            eng.edges_visible_batch_device(devA[k % NB], devB[k % NB], M, 0.01, dev_out)
        else:
            eng.feasible_batch_device(devQ[k % NB], M, dev_out)
            if with_dist:
                eng.distance_batch_device(devQ[k % NB], M, 0.5, False, dev_dist)

    def step_host(k):
        lib, h = eng.lib, eng.h
        from klampt_b200._capi import check
        import ctypes as C
        if edges:
            check(lib.kb_edges_visible_batch(h, C.c_void_p(hostA[k % NB].data_ptr()), C.c_void_p(hostB[k % NB].data_ptr()), M, 0.01, None,
                                             C.c_void_p(host_out.data_ptr()), None))
        else:
            check(lib.kb_feasible_batch(h, C.c_void_p(hostQ[k % NB].data_ptr()), M, C.c_void_p(host_out.data_ptr()), None))
            if with_dist:
                check(lib.kb_distance_batch(h, C.c_void_p(hostQ[k % NB].data_ptr()), M, 0.5, 0, C.c_void_p(host_dist.data_ptr()), None))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        with torch.cuda.stream(stream):
            for k in range(warm):
                fn(k)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(steps):
                fn(warm + k)
            e1.record(stream)
            barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput (value) with per-launch kernel timing for the roofline
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a few hundred ms before its first line: start it ahead of the warm-up
    eng.reset_stats()
    eng.set_option("time_kernels", 1)
    # warm-up first so the timed region's statistics are clean
    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            step_device(k)
    barrier()
    eng.reset_stats()
    t_clk0 = time.perf_counter()
    ms_dev = timed(step_device, args.steps, 0)
    st = eng.stats()
    eng.set_option("time_kernels", 0)
    # The timed region is a few tens of ms, shorter than nvidia-smi's sampling period: keep the same load running (untimed,
    # not counted) until at least three clock samples were taken under it.
    extra = 0
    with torch.cuda.stream(stream):
        while sampler.count(t_clk0) < 3 and time.perf_counter() - t_clk0 < 2.0:
            step_device(extra); extra += 1
            if extra % 4 == 0:
                torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop(window=(t_clk0, time.perf_counter()))
    clocks["window"] = "timed region + %d untimed trailing steps of the same load" % extra
    value = world * M * args.steps / (ms_dev * 1e-3)
    feas_frac = (st["configs_feasible"] / max(1, st["configs_checked"])) if not edges else (st["edges_visible"] / max(1, st["edges_checked"]))
    launches = st["kernel_launches"]

    # ---- end to end through the C ABI with host buffers
    eng.reset_stats()
    t_wall = []
    for k in range(args.warmup):
        step_host(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_host(args.warmup + k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = world * M * args.steps / float(tt.item())
    h2d = (2 if edges or with_dist else 1) * M * L * 8
    d2h = M + (8 * M if with_dist else 0)

    # ---- the same through the fp32-row entry point (kb_feasible_batch_f32: half the upload, values widened on the device); an extra
    #      line item, not the headline: the reference's interface passes doubles
    e2e_f32 = None
    if not edges and not with_dist:
        import ctypes as C
        from klampt_b200._capi import check
        hostQf = [q.to(torch.float32).pin_memory() for q in hostQ]

        def step_host_f32(k):
            check(eng.lib.kb_feasible_batch_f32(eng.h, C.c_void_p(hostQf[k % NB].data_ptr()), M, C.c_void_p(host_out.data_ptr()), None))
        for k in range(args.warmup):
            step_host_f32(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            step_host_f32(args.warmup + k)
        torch.cuda.synchronize()
        tf = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        e2e_f32 = {"value": world * M * args.steps / float(tf.item()), "unit": unit, "h2d_bytes_per_step": M * L * 4, "d2h_bytes_per_step": M}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (traversal) from the oracle's canonical traversal counts:
    #      B(c) = 4L + 1/8 + 32 n_box + 64 n_node + 72 n_tri + 16 n_pt   [bytes per configuration]
    from oracle.oracle import OracleWorld
    if orc is None:
        orc = OracleWorld(spec)
    peak, peak_src = peaks()
    ns = 20000
    Qs = gen_configs(0, ns)
    _, cnt = orc.feasible_batch(Qs, nthreads=NTHREADS, want_counts=True)
    bytes_per_cfg = 4 * L + 0.125 + 32 * cnt["n_box"].mean() + 64 * cnt["n_node"].mean() + 72 * cnt["n_tri"].mean() + 16 * cnt["n_pt"].mean()
    roofline = None
    # traverse_ms / traverse_launches were captured in `st` (device-resident run) before the e2e leg reset the statistics
    tl, tms = st["traverse_launches"], st["traverse_ms"]
    if tl > 0 and not edges:
        per_launch_cfg = M * args.steps / tl
        avg_ms = tms / tl
        achieved = bytes_per_cfg * per_launch_cfg / (avg_ms * 1e-3) / 1e9
        traffic, ncu_extra = None, None
        try:   # DRAM bytes of one launch from the committed ncu --set full capture, if it was taken on this launch shape
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("workload") == args.workload and abs(tj["configs_per_launch"] - per_launch_cfg) < 1:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                ncu_extra = {k: tj[k] for k in ("issue_active_pct", "l2_gbs", "dram_gbs", "warp_instructions") if k in tj}
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "ncu": ncu_extra, "kernel": "kb_traverse_kernel<0>", "algorithmic_bytes_per_config": bytes_per_cfg, "configs_per_launch": per_launch_cfg,
                    "avg_launch_ms": avg_ms, "kernel_share_of_step": tms / ms_dev, "peak_source": peak_src,
                    "counts_per_config": {k: float(cnt[k].mean()) for k in cnt.dtype.names}}

    # ---- CPU baseline on this box's host cores (bounded sample; rank 0 at N=1 only)
    cpu_baseline = None
    if world == 1:
        if edges:
            A0, B0 = synth.sample_edges(robot, lambda Q: eng.feasible_batch(Q), 4000, 99)
            cv, cores, cdone, cused, cfeas = cpu_leg(orc, lambda k, n: (A0, B0), args.cpu_seconds, edges=True)
        else:
            cv, cores, cdone, cused, cfeas = cpu_leg(orc, gen_configs, args.cpu_seconds, slice_n=(10_000 if with_dist else 100_000), with_dist=with_dist)
        t1 = time.perf_counter()
        n1 = 20000 if not edges else 500
        if edges:
            orc.edges_visible_batch(A0[:n1], B0[:n1], eps=0.01, nthreads=1)
        else:
            orc.feasible_batch(gen_configs(0, n1), nthreads=1)
            if with_dist:
                orc.distance_batch(gen_configs(0, n1), upper_bound=0.5, include_self=False, nthreads=1)
        single = n1 / (time.perf_counter() - t1)
        cpu_baseline = {"value": cv, "unit": unit, "cores": cores, "kind": "port",
                        "sample": "%d %s in %.1f s (successive slices of the same workload), oracle + OpenMP on all host threads; single thread: %.0f %s"
                                  % (cdone, "edges" if edges else "configurations", cused, single, unit),
                        "single_core": single, "feasible_fraction": cfeas}

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 traversal + f64 FK/recheck",
            "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_per_step": M, "links": L, "triangles": spec.total_tris(), "feasible_fraction": feas_frac,
                       "l2_policy": "inputs rotate over %d batches (%d MB) and static BVH data is %d MB, both larger than L2"
                                    % (NB, NB * M * L * 8 >> 20, eng.layout()["static_bytes"] >> 20),
                       "parallelism": "configs sharded, geometry replicated" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_f32_rows": e2e_f32, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
