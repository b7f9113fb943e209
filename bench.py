#!/usr/bin/env python
"""bench.py -- collision-checked configurations / second and edge checks / second on B200 for the batched
configuration-feasibility path.

  python bench.py [--gpus N --steps K --warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5] [--configs M] [--extras 0|1]

A "step" is one pass of the hot path (FK -> limits -> environment + self collision [-> distance]) over one batch of
synthetic configurations (or edges) of the workload.

Headline line (one JSON line on stdout, rank 0): BASELINE.json configs[1] -- the 6-DOF arm in the cluttered mesh world
(200 obstacles, ~500k triangles), 1M uniform random configurations per GPU per step.  Under torchrun (N > 1) every rank holds a
replica of the static geometry and checks its own 1M configurations: weak scaling, no data-path collective.

  value         configs/s, inputs already resident in HBM when the timed region starts (kb_feasible_batch_device)
  e2e           the same metric through the C ABI host entry point (kb_feasible_batch) with pinned host buffers:
                H2D of the configurations and D2H of the results are inside the timed region
  roofline      dominant kernel: algorithmic bytes per launch (counts of the oracle's canonical traversal) /
                measured launch duration vs the measured HBM peak
  cpu_baseline  the CPU restatement of the reference path on this box's host cores (bounded sample, rank 0 at N = 1):
                the SAH / -march=native build of the oracle ("fast"), with the strict checker build beside it
  strong        (N > 1) ONE 1M batch split with shard_range over the ranks, each shard through kb_feasible_batch_bits,
                result bitmasks all-gathered over NCCL inside the timer
  extras        the other BASELINE.json configs at their stated sizes, each with value / e2e / roofline / cpu_baseline:
                C1 10k configurations, C3 10M configurations (total, sharded over the ranks), C4 1M edges (total, block-cyclic
                over the ranks), C5 1M configurations collide + distance vs the 5M-point cloud (total, sharded).  Strong
                scaling: the totals are fixed, results are gathered as bitmasks (+ distances) inside the e2e timer.
                C6 (no BASELINE config): depth images of the C2 world, 640 x 480 rays each, through kb_raycast_batch.

--impl reference times the CPU path alone (rank 0) on the same workload and the same per-step batch size.  The real reference
cannot be built here (its arithmetic lives in the absent KrisLibrary), so this is the oracle port with all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
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
SUB = 1_000_000          # configurations (edges) per call into the engine: larger jobs loop over sub-batches of this size
NB = 4                   # rotating input batches per job (together larger than L2)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--configs", type=int, default=0, help="configurations (or edges) per step: per GPU for the headline, total for extras")
    ap.add_argument("--extras", type=int, default=1, help="1: also measure the other BASELINE configs and the ray-cast workload (reported under 'extras'); 2: the ray-cast workload only")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the headline cpu_baseline leg (extras: a third of it)")
    return ap.parse_args()


WORKLOADS = {
    "c1": dict(name="C1 arm6 + ground + 10 boxes (self+env), uniform random configurations", n=10_000, kind="feas"),
    "c2": dict(name="C2 arm6 (TX90-class, 7 links) in cluttered mesh world: 200 blob obstacles, ~500k triangles, self+env, "
                    "uniform random configurations", n=1_000_000, kind="feas"),
    "c3": dict(name="C3 dual-arm 15-DOF (18 links) self-collision only, uniform random configurations", n=10_000_000, kind="feas"),
    "c4": dict(name="C4 straight-line edges in the C2 world at eps=0.01 (EpsilonEdgeChecker), early exit", n=1_000_000, kind="edges"),
    "c5": dict(name="C5 arm6 meshes vs 5M-point cloud (points on the C2 obstacle surfaces + 5 mm noise, margin 5 mm): collide bit + "
                    "min distance with upperBound 0.5 m per configuration", n=1_000_000, kind="dist"),
}
_WORLD_CACHE = {}


def make_world(which):
    from klampt_b200 import synth
    key = {"c4": "c2"}.get(which, which)
    if key not in _WORLD_CACHE:
        _WORLD_CACHE[key] = {"c1": synth.world_c1, "c2": synth.world_c2, "c3": synth.world_c3, "c5": synth.world_c5}[key]()
    return _WORLD_CACHE[key]


def config_dict(which, per_step, spec):
    """identical on both arms (ours / reference) so the driver can tell they ran the same thing"""
    L = spec.robot.L
    return {"workload": WORKLOADS[which]["name"], "per_gpu_per_step": int(per_step), "links": int(L), "triangles": int(spec.total_tris()),
            "l2_policy": "inputs rotate over %d distinct batches (%d MB) and the static BVH data (C2: 180 MB) is larger than L2 as well"
                         % (NB, NB * min(per_step, SUB) * L * 8 * (2 if WORKLOADS[which]["kind"] == "edges" else 1) >> 20)}


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


# ------------------------------------------------------------------------------------------------------- CPU arm
def cpu_run(orc, kind, data):
    if kind == "edges":
        out, _ = orc.edges_visible_batch(data[0], data[1], eps=0.01, nthreads=NTHREADS)
    else:
        out = orc.feasible_batch(data, nthreads=NTHREADS)
        if kind == "dist":
            orc.distance_batch(data, upper_bound=0.5, include_self=False, nthreads=NTHREADS)
    return out


def cpu_leg(orc, kind, gen_batch, budget_s, slice_n):
    """oracle with all host threads on successive slices of the workload until the time budget is spent"""
    done, t_used, ones, k = 0, 0.0, 0, 0
    while (t_used < budget_s or k == 0) and k < 64:
        data = gen_batch(k, slice_n)
        t0 = time.perf_counter()
        out = cpu_run(orc, kind, data)
        t_used += time.perf_counter() - t0
        done += len(out)
        ones += int(out.sum())
        k += 1
    return done / t_used, done, t_used, ones / max(1, done)


def cpu_baseline_for(which, spec, gen_batch, budget_s, unit, strict):
    """the fast (SAH, -march=native, FMA) build of the oracle on all host threads + one thread, and the strict checker build"""
    from oracle.oracle import OracleWorld
    kind = WORKLOADS[which]["kind"]
    slice_n = {"feas": 100_000, "edges": 2_000, "dist": 4_000}[kind]
    if which == "c1":
        slice_n = 10_000
    fast = OracleWorld(spec, variant="fast")
    v, done, used, frac = cpu_leg(fast, kind, gen_batch, budget_s, slice_n)
    n1 = max(64, slice_n // 16)
    d1 = gen_batch(0, n1)
    d1 = (d1[0][:n1], d1[1][:n1]) if kind == "edges" else d1[:n1]
    t1 = time.perf_counter()
    if kind == "edges":
        fast.edges_visible_batch(d1[0], d1[1], eps=0.01, nthreads=1)
    else:
        fast.feasible_batch(d1, nthreads=1)
        if kind == "dist":
            fast.distance_batch(d1, upper_bound=0.5, include_self=False, nthreads=1)
    single = n1 / (time.perf_counter() - t1)
    fast.close()
    sv, sdone, sused, _ = cpu_leg(strict, kind, gen_batch, max(1.0, budget_s / 3), slice_n)
    what = "edges" if kind == "edges" else "configurations"
    return {"value": v, "unit": unit, "cores": NTHREADS, "kind": "port",
            "sample": "%d %s in %.1f s (successive slices of the same workload); oracle port, binned-SAH tree, gcc -O3 -march=native with FMA, "
                      "OpenMP over all %d host threads (the real reference is single-threaded)" % (done, what, used, NTHREADS),
            "single_core": single, "ones_fraction": frac,
            "strict_build": {"value": sv, "what": "the checker build (median-split tree, -march=x86-64-v3, no FMA contraction), %d %s in %.1f s" % (sdone, what, sused)}}


def canonical_bytes(which, orc, sample, L):
    """B = 4L + 1/8 + 32 n_box + 64 n_node + 72 n_tri + 16 n_pt per configuration (counts of the oracle's canonical
    traversal); per edge: the sum over the midpoints the sequential checker visits; C5 adds the canonical branch-and-bound
    distance traversal (+4 B distance out)"""
    from oracle.oracle import Counts
    kind = WORKLOADS[which]["kind"]

    def B(c):
        return 32.0 * c["n_box"] + 64.0 * c["n_node"] + 72.0 * c["n_tri"] + 16.0 * c["n_pt"]
    if kind == "edges":
        A, Bq = sample
        tot = {"n_box": 0, "n_node": 0, "n_tri": 0, "n_pt": 0}
        nchk = 0
        dp = C.POINTER(C.c_double)
        for i in range(len(A)):
            cnt = Counts()
            n = C.c_int32(0)
            a, b = np.ascontiguousarray(A[i]), np.ascontiguousarray(Bq[i])
            orc.L.ko_edge_visible(orc.h, a.ctypes.data_as(dp), b.ctypes.data_as(dp), 0.01, None, C.byref(n), C.byref(cnt))
            nchk += n.value
            for k in tot:
                tot[k] += getattr(cnt, k)
        per = {k: v / len(A) for k, v in tot.items()}
        per["checks_per_edge"] = nchk / len(A)
        return 2 * 4 * L + 0.125 + (4 * L) * per["checks_per_edge"] + B(per), per
    _, cnt = orc.feasible_batch(sample, nthreads=NTHREADS, want_counts=True)
    per = {k: float(cnt[k].mean()) for k in cnt.dtype.names}
    total = 4 * L + 0.125 + B(per)
    if kind == "dist":
        dp = C.POINTER(C.c_double)
        tot = {"n_box": 0, "n_node": 0, "n_tri": 0, "n_pt": 0}
        m = min(len(sample), 400)
        for i in range(m):
            c2 = Counts()
            q = np.ascontiguousarray(sample[i])
            orc.L.ko_distance(orc.h, q.ctypes.data_as(dp), 0.5, 0, None, C.byref(c2))
            for k in tot:
                tot[k] += getattr(c2, k)
        dper = {k: v / m for k, v in tot.items()}
        total += 4 + B(dper)
        per.update({"dist_" + k: v for k, v in dper.items()})
    return total, per


# ------------------------------------------------------------------------------------------------------- reference arm
def reference_arm(args):
    from klampt_b200 import synth
    from oracle.oracle import OracleWorld
    which = args.workload
    wl = WORKLOADS[which]
    kind = wl["kind"]
    spec = make_world(which)
    M = args.configs or (1_000_000 if which in ("c2", "c3") else wl["n"])
    edges = kind == "edges"
    unit = "edges/s" if edges else "configs/s"
    # each step is the SAME per-step batch our arm runs when that stays within a few minutes for W + K steps; edges and distance
    # queries are three to four orders of magnitude slower per unit on the CPU, so their steps are bounded samples
    per_step = M if kind == "feas" else min(M, 4_000 if edges else 8_000)
    orc = OracleWorld(spec, variant="fast")
    if edges:
        A0, B0 = synth.sample_edges(spec.robot, lambda Q: orc.feasible_batch(Q, nthreads=NTHREADS), per_step, 4)
        gen = lambda k, n: (A0, B0)
    else:
        gen = lambda k, n: synth.sample_configs(spec.robot, n, 1000 + 17 * (k % NB))
    times = []
    for s in range(args.warmup + args.steps):
        data = gen(s, per_step)
        t0 = time.perf_counter()
        cpu_run(orc, kind, data)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    v = per_step * len(times) / sum(times)
    line = {"impl": "reference", "metric": "edge checks/sec" if edges else "collision-checked configs/sec", "value": v, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(which, M, spec),
            "cpu_baseline": {"value": v, "unit": unit, "cores": NTHREADS, "kind": "port",
                             "sample": "%d %s per step%s; oracle port (binned-SAH tree, gcc -O3 -march=native with FMA), OpenMP over all %d host threads"
                                       % (per_step, "edges" if edges else "configurations",
                                          "" if per_step == M else " (bounded sample of the %d-per-step workload)" % M, NTHREADS)},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------- our arm
class Job:
    """one workload on this rank's GPU: `n_local` units per step, processed in sub-batches of at most SUB through the engine"""

    def __init__(self, which, n_local, rank, world, local_rank, stream, seed_base=1000, options=None):
        import torch
        from klampt_b200 import synth
        from klampt_b200.engine import Engine
        self.torch = torch
        self.which, self.kind = which, WORKLOADS[which]["kind"]
        self.spec = make_world(which)
        self.robot = self.spec.robot
        self.L = self.robot.L
        self.rank, self.world = rank, world
        self.eng = Engine(self.spec, device=local_rank, options=options)
        self.eng.set_stream(stream.cuda_stream)
        self.stream = stream
        self.n_local = int(n_local)
        self.sub = min(SUB, self.n_local)
        self.nsub = -(-self.n_local // self.sub)
        sub, L = self.sub, self.L
        self.seed_base = seed_base
        if self.kind == "edges":
            self.hostA, self.hostB = [], []
            for k in range(NB):
                if k >= 2:       # two distinct edge batches (2 x 2 x sub x L x 8 B, larger than L2 at the bench size) are reused
                    self.hostA.append(self.hostA[k - 2]); self.hostB.append(self.hostB[k - 2])
                    continue
                A, B = synth.sample_edges(self.robot, lambda Q: self.eng.feasible_batch(Q), sub, 4 + k + 10 * rank)
                self.hostA.append(torch.from_numpy(A).pin_memory())
                self.hostB.append(torch.from_numpy(B).pin_memory())
            self.devA = [a.cuda(non_blocking=True) for a in self.hostA[:2]] * 2
            self.devB = [b.cuda(non_blocking=True) for b in self.hostB[:2]] * 2
        else:
            self.hostQ = [torch.from_numpy(self.gen_configs(k, sub)).pin_memory() for k in range(NB)]
            self.devQ = [q.cuda(non_blocking=True) for q in self.hostQ]
        self.host_out = torch.empty(sub, dtype=torch.uint8).pin_memory()
        self.dev_out = torch.empty(sub, dtype=torch.uint8, device="cuda")
        self.dev_bits = torch.zeros((sub + 31) // 32, dtype=torch.int32, device="cuda")
        self.host_bits = torch.zeros((sub + 7) // 8, dtype=torch.uint8).pin_memory()
        if self.kind == "dist":
            self.dev_dist = torch.empty(sub, dtype=torch.float64, device="cuda")
            self.host_dist = torch.empty(sub, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()

    def gen_configs(self, k, n):
        from klampt_b200 import synth
        return synth.sample_configs(self.robot, n, self.seed_base + 17 * (k % NB) + self.rank * 101)

    def sizes(self):
        for j in range(self.nsub):
            yield j, min(self.sub, self.n_local - j * self.sub)

    def step_device(self, k):
        e = self.eng
        for j, n in self.sizes():
            b = (k * self.nsub + j) % NB
            if self.kind == "edges":
                e.edges_visible_batch_device(self.devA[b], self.devB[b], n, 0.01, self.dev_out)
            else:
                e.feasible_batch_device(self.devQ[b], n, self.dev_out)
                if self.kind == "dist":
                    e.distance_batch_device(self.devQ[b], n, 0.5, False, self.dev_dist)

    def step_host(self, k, bits=False):
        from klampt_b200._capi import check
        lib, h = self.eng.lib, self.eng.h
        out = self.host_bits if bits else self.host_out
        for j, n in self.sizes():
            b = (k * self.nsub + j) % NB
            if self.kind == "edges":
                fn = lib.kb_edges_visible_batch_bits if bits else lib.kb_edges_visible_batch
                check(fn(h, C.c_void_p(self.hostA[b].data_ptr()), C.c_void_p(self.hostB[b].data_ptr()), n, 0.01, None, C.c_void_p(out.data_ptr()), None))
            else:
                if bits:
                    check(lib.kb_feasible_batch_bits(h, C.c_void_p(self.hostQ[b].data_ptr()), n, C.c_void_p(out.data_ptr())))
                else:
                    check(lib.kb_feasible_batch(h, C.c_void_p(self.hostQ[b].data_ptr()), n, C.c_void_p(out.data_ptr()), None))
                if self.kind == "dist":
                    check(lib.kb_distance_batch(h, C.c_void_p(self.hostQ[b].data_ptr()), n, 0.5, 0, C.c_void_p(self.host_dist.data_ptr()), None))

    def h2d_d2h(self, bits=False):
        per_unit_in = self.L * 8 * (2 if self.kind == "edges" else 1) * (2 if self.kind == "dist" else 1)
        per_unit_out = (0.125 if bits else 1) + (8 if self.kind == "dist" else 0)
        return int(self.n_local * per_unit_in), int(self.n_local * per_unit_out)


def main():
    args = parse()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            reference_arm(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    from klampt_b200 import synth
    from klampt_b200.shard import shard_range, interleaved_indices

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    peak, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device(fn, steps):
        """K steps on the engine's stream bracketed by barrier + synchronize, CUDA events, max over ranks -> ms"""
        with torch.cuda.stream(stream):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(steps):
                fn(k)
            e1.record(stream)
            barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def timed_host(fn, steps, warm):
        for k in range(warm):
            fn(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            fn(warm + k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt)

    def gather_bits(job, local_bits_host, n_total):
        """the ONE exchange step of the path: every rank's packed result bitmask, padded to the largest shard"""
        per = (-(-n_total // world) + 7) // 8
        t = torch.zeros(per, dtype=torch.uint8, device="cuda")
        t[:local_bits_host.numel()].copy_(local_bits_host[:per], non_blocking=True)
        full = torch.empty(per * world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(full, t)
        return full.cpu()

    def measure(which, n_local, n_total, steps, warm, headline, cpu_budget):
        """value / e2e / roofline (/ cpu_baseline on rank 0 at N = 1) of one workload"""
        wl = WORKLOADS[which]
        kind = wl["kind"]
        edges = kind == "edges"
        unit = "edges/s" if edges else "configs/s"
        job = Job(which, n_local, rank, world, local_rank, stream)
        eng = job.eng
        sampler = ClockSampler(local_rank) if headline else None
        if sampler:
            sampler.start()          # nvidia-smi needs a few hundred ms before its first line: start it ahead of the warm-up
        eng.set_option("time_kernels", 1)
        with torch.cuda.stream(stream):
            for k in range(warm):
                job.step_device(k)
        barrier()
        eng.reset_stats()
        t_clk0 = time.perf_counter()
        ms_dev = timed_device(lambda k: job.step_device(warm + k), steps)
        st = eng.stats()
        eng.set_option("time_kernels", 0)
        clocks = None
        if sampler:
            # the timed region can be shorter than nvidia-smi's sampling period: keep the same load running (untimed, not counted)
            # until at least three clock samples were taken under it
            extra = 0
            with torch.cuda.stream(stream):
                while sampler.count(t_clk0) < 3 and time.perf_counter() - t_clk0 < 2.0:
                    job.step_device(extra); extra += 1
                    if extra % 4 == 0:
                        torch.cuda.synchronize()
            torch.cuda.synchronize()
            clocks = sampler.stop(window=(t_clk0, time.perf_counter()))
            clocks["window"] = "timed region + %d untimed trailing steps of the same load" % extra
        units_all = n_total if n_total else world * n_local
        value = units_all * steps / (ms_dev * 1e-3)
        ones = (st["edges_visible"] / max(1, st["edges_checked"])) if edges else (st["configs_feasible"] / max(1, st["configs_checked"]))

        # ---- end to end through the C ABI with pinned host buffers (+ the NCCL gather of result bitmasks when the job is sharded)
        eng.reset_stats()
        sharded = world > 1 and n_total
        if sharded:
            def e2e_step(k):
                job.step_host(k, bits=True)
                gather_bits(job, job.host_bits, n_total)
        else:
            e2e_step = lambda k: job.step_host(k)
        dt = timed_host(e2e_step, steps, warm)
        h2d, d2h = job.h2d_d2h(bits=bool(sharded))
        e2e = {"value": units_all * steps / dt, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        if sharded:
            e2e["gather"] = "NCCL all_gather_into_tensor of %d-byte result bitmasks inside the timer" % ((-(-n_total // world) + 7) // 8)

        e2e_f32 = None
        if headline and kind == "feas":
            # the fp32-row entry point (half the upload, values widened on the device): an extra item, not the headline
            from klampt_b200._capi import check
            hostQf = [q.to(torch.float32).pin_memory() for q in job.hostQ]

            def step_f32(k):
                check(eng.lib.kb_feasible_batch_f32(eng.h, C.c_void_p(hostQf[k % NB].data_ptr()), job.sub, C.c_void_p(job.host_out.data_ptr()), None))
            tf = timed_host(step_f32, steps, warm)
            e2e_f32 = {"value": world * job.sub * steps / tf, "unit": unit, "h2d_bytes_per_step": job.sub * job.L * 4, "d2h_bytes_per_step": job.sub}

        res = {"workload": wl["name"], "metric": "edge checks/sec" if edges else "collision-checked configs/sec", "value": value, "unit": unit,
               "ms_per_step": ms_dev / steps, "steps": steps, "warmup": warm, "units_per_step_all_gpus": int(units_all),
               "scaling": "strong" if n_total else "weak", "e2e": e2e, "ones_fraction": ones, "gpu_launches": int(st["kernel_launches"])}
        if edges:
            res["config_checks_per_edge"] = st["edge_config_checks"] / max(1, st["edges_checked"])
            res["config_checks_per_s"] = value * res["config_checks_per_edge"]
        if rank == 0:
            # ---- roofline of the dominant kernel from the oracle's canonical traversal counts
            if edges:
                sample = (job.hostA[0][:600].numpy(), job.hostB[0][:600].numpy())
            else:
                sample = job.gen_configs(0, 20000 if kind == "feas" else 4000)
            from oracle.oracle import OracleWorld
            strict = OracleWorld(job.spec)        # the checker build: canonical median-split tree (roofline counts + the strict CPU leg)
            bytes_per_unit, counts = canonical_bytes(which, strict, sample, job.L)
            tl, tms = st["traverse_launches"], st["traverse_ms"]
            if tl > 0:
                units_rank = n_local * steps
                achieved = bytes_per_unit * units_rank / (tms * 1e-3) / 1e9
                traffic, ncu_extra = None, None
                try:   # DRAM bytes of one launch from the committed ncu --set full capture, if it was taken on this launch shape
                    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                    if tj.get("workload") == which and abs(tj["configs_per_launch"] - units_rank / tl) < 1:
                        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                        ncu_extra = {k: tj[k] for k in ("issue_active_pct", "l2_gbs", "dram_gbs", "warp_instructions", "capture") if k in tj}
                except Exception:
                    pass
                res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                                   "ncu": ncu_extra, "kernel": {"feas": "kb_traverse_wide_kernel", "edges": "kb_traverse_wide_kernel (edge midpoints)",
                                                                "dist": "kb_traverse_kernel + kb_distance_kernel"}[kind],
                                   "algorithmic_bytes_per_unit": bytes_per_unit, "units_per_launch": units_rank / tl, "avg_launch_ms": tms / tl,
                                   "kernel_share_of_step": tms / ms_dev, "peak_source": peak_src, "counts_per_unit": counts}
            if world == 1:
                if edges:
                    A0, B0 = job.hostA[0].numpy(), job.hostB[0].numpy()
                    gen = lambda k, n: (A0[(k * n) % (len(A0) - n):][:n], B0[(k * n) % (len(A0) - n):][:n])
                else:
                    gen = job.gen_configs
                res["cpu_baseline"] = cpu_baseline_for(which, job.spec, gen, cpu_budget, unit, strict)
            strict.close()
        res["_feas_frac"], res["_clocks"], res["_e2e_f32"], res["_static_mb"] = ones, clocks, e2e_f32, eng.layout()["static_bytes"] >> 20
        job.eng.close()
        del job
        torch.cuda.empty_cache()
        return res


    def measure_rays(steps, warm, cpu_budget):
        """the 'next' row 8f-4 as a workload: depth images of the C2 world from a camera circling it (CameraSensor's ray-cast path,
        VisualSensors.cpp:413-475): 640 x 480 rays per image, one launch per image, 4 images per step; every rank renders its own
        views (weak scaling).  value: rays device-resident; e2e: kb_raycast_batch with pinned host rays and results."""
        from klampt_b200 import sensing
        from klampt_b200.engine import Engine
        from klampt_b200._capi import check
        spec = make_world("c2")
        eng = Engine(spec, device=local_rank)
        eng.set_stream(stream.cuda_stream)
        q = synth.sample_configs(spec.robot, 1, 77)[0]
        qc = np.ascontiguousarray(q)
        per_step, W, H = 4, 640, 480
        host, dev, cams = [], [], []
        for k in range(8):
            a = 2.0 * math.pi * (k + 8 * rank) / (8 * world)
            eye = np.array([3.2 * math.cos(a), 3.2 * math.sin(a), 1.3])
            fwd = np.array([0.0, 0.0, 0.5]) - eye
            fwd /= np.linalg.norm(fwd)
            right = np.cross(fwd, [0.0, 0.0, 1.0]); right /= np.linalg.norm(right)
            down = np.cross(fwd, right)
            cam = sensing.CameraSensor(W, H, zmin=0.1, zmax=8.0, Tsensor=synth.make_T(np.stack([right, down, fwd], axis=1), eye))
            cams.append(cam)
            r, _, _ = cam.rays()
            host.append(torch.from_numpy(np.ascontiguousarray(r)).pin_memory())
            dev.append(host[-1].cuda(non_blocking=True))
        n = W * H
        d_id, d_dist = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
        h_id, h_dist = torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
        qp = C.c_void_p(qc.ctypes.data)

        def step_device(k):
            for j in range(per_step):
                check(eng.lib.kb_raycast_batch_device(eng.h, qp, C.c_void_p(dev[(k * per_step + j) % 8].data_ptr()), n, None,
                                                      C.c_void_p(d_id.data_ptr()), C.c_void_p(d_dist.data_ptr()), None))

        def step_host(k):
            for j in range(per_step):
                check(eng.lib.kb_raycast_batch(eng.h, qp, C.c_void_p(host[(k * per_step + j) % 8].data_ptr()), n, None,
                                               C.c_void_p(h_id.data_ptr()), C.c_void_p(h_dist.data_ptr()), None))
        from klampt_b200._capi import KbCamera
        kcams = []
        for cam in cams:
            kc = KbCamera()
            kc.pose[:] = list(cam.Tsensor)
            kc.fx, kc.fy, kc.cx, kc.cy = cam.viewport()
            kc.zmin, kc.zmax, kc.xres, kc.yres = cam.zmin, cam.zmax, W, H
            kcams.append(kc)
        h_depth = torch.empty(n, dtype=torch.float32).pin_memory()

        def step_camera(k):
            for j in range(per_step):
                check(eng.lib.kb_camera_depth(eng.h, qp, C.byref(kcams[(k * per_step + j) % 8]), None, C.c_void_p(h_depth.data_ptr()), C.c_void_p(h_id.data_ptr())))
        with torch.cuda.stream(stream):
            for k in range(warm):
                step_device(k)
        barrier()
        eng.reset_stats()
        ms_dev = timed_device(lambda k: step_device(warm + k), steps)
        launches = int(eng.stats()["kernel_launches"])
        units_all = world * per_step * n
        value = units_all * steps / (ms_dev * 1e-3)
        dt = timed_host(step_host, steps, warm)
        dt_cam = timed_host(step_camera, steps, warm)
        hit_frac = float((h_id >= 0).float().mean())
        res = {"workload": "C6 depth images of the C2 world (arm6 at one configuration + 200 blob obstacles, ~500k triangles): 640 x 480 rays per image "
                           "from a camera circling the scene, nearest hit + world id per ray (WorldModel::RayCast per pixel)",
               "metric": "rays/sec", "value": value, "unit": "rays/s", "ms_per_step": ms_dev / steps, "steps": steps, "warmup": warm,
               "units_per_step_all_gpus": int(units_all), "scaling": "weak",
               "e2e": {"value": units_all * steps / dt, "unit": "rays/s", "h2d_bytes_per_step": per_step * n * 48, "d2h_bytes_per_step": per_step * n * 12},
               "e2e_camera": {"value": units_all * steps / dt_cam, "unit": "rays/s", "h2d_bytes_per_step": per_step * 176, "d2h_bytes_per_step": per_step * n * 8,
                              "what": "kb_camera_depth: the rays are built on the device from the camera's pose and intrinsics; float depth + world id per pixel come back"},
               "ones_fraction": hit_frac, "gpu_launches": launches,
               "l2_policy": "8 distinct images (118 MB of rays) rotate; the static data (258 MB) is larger than L2"}
        if rank == 0:
            from oracle.oracle import OracleWorld
            strict = OracleWorld(spec)
            sample = host[0].numpy()[:: 61]
            cnt = strict.raycast_counts(q, sample)
            bytes_per_ray = 48 + 12 + 32.0 * cnt["n_node"] + 72.0 * cnt["n_tri"] + 32.0 * cnt["n_pt"]
            achieved = bytes_per_ray * (per_step * n * steps) / (ms_dev * 1e-3) / 1e9
            res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                               "kernel": "kb_raycast_kernel", "algorithmic_bytes_per_unit": bytes_per_ray, "units_per_launch": n,
                               "avg_launch_ms": ms_dev / steps / per_step, "kernel_share_of_step": None, "peak_source": peak_src,
                               "counts_per_unit": cnt,
                               "note": "bytes of the reference's per-body loop (every body's root box + its hierarchy: 32 B per box test, 72 B per "
                                       "triangle); the kernel visits far fewer boxes because the bodies sit under a top-level hierarchy, so the "
                                       "fraction can exceed 1 -- it compares work done per second, not DRAM traffic"}
            strict.close()
            if world == 1:
                fast = OracleWorld(spec, variant="fast")
                done, used, k = 0, 0.0, 0
                while (used < cpu_budget or k == 0) and k < 64:
                    r = host[k % 8].numpy()[(k // 8) % 4:: 4][:60000]
                    t0 = time.perf_counter()
                    fast.raycast_batch(q, r, nthreads=NTHREADS)
                    used += time.perf_counter() - t0
                    done += len(r)
                    k += 1
                fast.close()
                res["cpu_baseline"] = {"value": done / used, "unit": "rays/s", "cores": NTHREADS, "kind": "port",
                                       "sample": "%d rays of the same images in %.1f s; oracle port (per-body loop as WorldModel::RayCast, binned-SAH "
                                                 "trees, gcc -O3 -march=native), OpenMP over all %d host threads" % (done, used, NTHREADS)}
        eng.close()
        torch.cuda.empty_cache()
        return res

    # ------------------------------------------------------------------------------------------ headline
    which = args.workload
    wl = WORKLOADS[which]
    M = args.configs or (1_000_000 if which in ("c2", "c3") else wl["n"])
    head = measure(which, M, 0, args.steps, args.warmup, True, args.cpu_seconds)

    # ------------------------------------------------------------------------------------------ strong-scaling leg of the headline
    strong = None
    if world > 1 and wl["kind"] == "feas":
        total = M
        lo, hi = shard_range(total, rank, world)
        job = Job(which, hi - lo, 0, world, local_rank, stream)       # rank 0's batch on every rank: the same seed, then this rank's block
        full = [job.gen_configs(k, total) for k in range(2)]
        for k in range(NB):
            job.hostQ[k].copy_(torch.from_numpy(full[k % 2][lo:hi]))
        # the 1-GPU answer of the same batch, for the equality check below (computed once, outside the timers)
        def strong_step(k):
            job.step_host(k, bits=True)
            return gather_bits(job, job.host_bits, total)
        dt = timed_host(strong_step, args.steps, args.warmup)
        got = strong_step(0).numpy()
        per = (-(-total // world) + 7) // 8
        bits = np.concatenate([np.unpackbits(got[r * per:(r + 1) * per], bitorder="little")[:shard_range(total, r, world)[1] - shard_range(total, r, world)[0]] for r in range(world)])
        ok = None
        if rank == 0:
            mine = job.eng.feasible_batch(full[0][lo:hi])
            ok = bool(np.array_equal(bits[lo:hi], mine)) and len(bits) == total
        strong = {"what": "ONE batch of %d configurations split with shard_range over %d ranks, kb_feasible_batch_bits per shard (pinned host buffers), "
                          "NCCL all-gather of the result bitmasks inside the timer" % (total, world),
                  "value": total * args.steps / dt, "unit": "configs/s", "ms_per_step": 1e3 * dt / args.steps, "scaling": "strong",
                  "gathered_equals_local": ok, "feasible_fraction": float(bits.mean())}
        job.eng.close()
        del job

    # ------------------------------------------------------------------------------------------ the other BASELINE configs
    extras = []
    if args.extras and which == "c2":
        xs, xw = max(3, args.steps // 3), 3
        for w2 in (("c1", "c3", "c4", "c5") if args.extras == 1 else ()):      # --extras 2: the ray workload only
            total = WORKLOADS[w2]["n"]
            if w2 == "c1":
                if rank == 0 and world == 1:
                    extras.append(measure(w2, total, 0, max(10, args.steps), xw, False, args.cpu_seconds / 3))
                continue
            if WORKLOADS[w2]["kind"] == "edges":
                n_local = len(interleaved_indices(total, rank, world))
            else:
                lo, hi = shard_range(total, rank, world)
                n_local = hi - lo
            extras.append(measure(w2, n_local, total, xs, xw, False, args.cpu_seconds / 3))
        extras.append(measure_rays(max(10, args.steps), xw, args.cpu_seconds / 3))

    if rank == 0:
        spec = make_world(which)
        for r in extras:
            for k in [k for k in r if k.startswith("_")]:
                r.pop(k)
        cfg = config_dict(which, M, spec)
        line = {"metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 traversal + f64 FK/recheck", "data": "synthetic", "config": cfg,
                "feasible_fraction": head["_feas_frac"], "static_data_mb": head["_static_mb"],
                "parallelism": "configs sharded, geometry replicated" if world > 1 else "single GPU",
                "e2e": head["e2e"], "e2e_f32_rows": head["_e2e_f32"], "gpu_launches": head["gpu_launches"], "clocks": head["_clocks"],
                "roofline": head.get("roofline"), "cpu_baseline": head.get("cpu_baseline"), "strong": strong, "extras": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
