#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 backend (BASELINE.json metric:
"LM iters/sec (10k-KF graph) & RANSAC Mpts/sec").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one GraphSLAM::optimize(20) pass over the cfg2 graph (10 000 keyframes, ~2 000 landmarks,
~60 000 edges; BASELINE.json configs[1]) from the same initial estimates, followed by one RANSAC
plane-fit pass over the cfg3 frame (640x480 cloud, 64 bbox crops x 1024 hypotheses; configs[2]).
The JSON line's `value` is LM iterations per second with all inputs resident in HBM; `e2e` is the
same metric through the C-ABI with host buffers (CSR build + H2D + LM loop + D2H inside the timed
region).  The RANSAC numbers ride in the `ransac` object of the same line.
With N > 1 GPUs the SAME graph is sharded by contiguous keyframe range over the N ranks (strong scaling: one graph,
csrc/ssb_peer.cuh) and the 64 crops of the frame are dealt round-robin to the ranks; `value` is that one job's rate.
`--impl reference` times the CPU restatement of the reference's g2o / PCL path (oracle/, "port":
/root/reference itself cannot be compiled here) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

LM_ITERS = 20
_MARG_CPU_NOTE = ("oracle, one sparse Cholesky of the full system + g2o's MarginalCovarianceCholesky recursion (memoised "
                  "elements of the inverse; what computeMarginals runs in the reference), 1 thread")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (arrival time, csv line)
        self.t_mark = None

    def mark(self):
        """Start of the timed region: only samples that arrive after this (and before stop()) are reported.  The
        nvidia-smi process itself is started earlier (its start-up takes longer than a short timed region)."""
        self.t_mark = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.monotonic(), ln.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_stop = time.monotonic()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t_mark if self.t_mark is not None else 0.0
        for t_ln, ln in self.lines:
            if t_ln < t0 or t_ln > t_stop:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def graph_bytes(spec):
    Np, Nl = spec.n_poses, spec.n_landmarks
    El = int((spec.ekind == 1).sum())
    Epp = int((spec.ekind == 0).sum())
    # algorithmic bytes of one matrix-free Schur PCG iteration
    b_cg = Np * 288 + Epp * 288 + 2 * El * 144 + Nl * 72 + Np * 288 + 10 * Np * 48 + 2 * Nl * 24
    b_lin = (Epp * 232 + El * 80 + Np * 56 + Nl * 24) + (Np * (288 + 48) + Epp * 288 + Nl * (72 + 24) + El * 144)
    b_chi2 = Epp * 232 + El * 80 + Np * 56 + Nl * 24 + 8
    b_upd = (Np * 48 + Nl * 24) + 2 * (Np * 56 + Nl * 24)
    h2d = Np * 64 + Nl * 32 + El * 80 + Epp * 240 + (Nl + 1 + 2 * (Np + 1) + 2 * El + 2 * Epp) * 4 + Np + Nl
    d2h = Np * 64 + Nl * 32
    return dict(Np=Np, Nl=Nl, El=El, Epp=Epp, b_cg=b_cg, b_lin=b_lin, b_chi2=b_chi2, b_upd=b_upd, h2d=h2d, d2h=d2h)


class FrameRecorder:
    """GraphSLAM proxy for the per-frame loop: remembers the add_* calls and, at chosen frames, the estimates right before
    optimize(), so that exactly that frame's optimize() can be repeated on the CPU oracle afterwards (same graph, same
    starting estimates) and compared in time and in result."""

    def __init__(self, graph, probe_frames):
        self.g = graph
        self.calls = []
        self.kinds = []                      # per vertex id: 0 = SE3, 1 = XYZ
        self.probe_frames = set(probe_frames)
        self.n_opt = 0
        self.probes = []

    def __getattr__(self, name):
        return getattr(self.g, name)

    def add_se3_node(self, T):
        self.calls.append(("add_se3_node", (np.array(T, dtype=np.float64),)))
        self.kinds.append(0)
        return self.g.add_se3_node(T)

    def add_point_xyz_node(self, x):
        self.calls.append(("add_point_xyz_node", (np.array(x, dtype=np.float64),)))
        self.kinds.append(1)
        return self.g.add_point_xyz_node(x)

    def add_se3_edge(self, a, b, Z, info):
        self.calls.append(("add_se3_edge", (a, b, np.array(Z, dtype=np.float64), np.array(info, dtype=np.float64))))
        return self.g.add_se3_edge(a, b, Z, info)

    def add_se3_point_xyz_edge(self, a, b, z, info):
        self.calls.append(("add_se3_point_xyz_edge", (a, b, np.array(z, dtype=np.float64), np.array(info, dtype=np.float64))))
        return self.g.add_se3_point_xyz_edge(a, b, z, info)

    def _estimates(self):
        return [self.g.get_se3(v) if k == 0 else self.g.get_point_xyz(v) for v, k in enumerate(self.kinds)]

    def optimize(self, max_iterations=1024):
        self.n_opt += 1
        if self.n_opt not in self.probe_frames:
            return self.g.optimize(max_iterations)
        before = self._estimates()
        t0 = time.perf_counter()
        r = self.g.optimize(max_iterations)
        dt = time.perf_counter() - t0
        self.probes.append({"frame": self.n_opt, "n_calls": len(self.calls), "n_vertices": len(self.kinds), "before": before,
                            "after": self._estimates(), "gpu_ms": 1e3 * dt, "max_iterations": max_iterations,
                            "lm_iterations": int(self.g.iterations), "trials": int((getattr(self.g, "stats", None) or {}).get("total_trials", -1))})
        return r

    def replay_on(self, oracle_graph, probe):
        for name, a in self.calls[:probe["n_calls"]]:
            getattr(oracle_graph, name)(*a)
        for v in range(probe["n_vertices"]):
            (oracle_graph.set_se3 if self.kinds[v] == 0 else oracle_graph.set_point_xyz)(v, probe["before"][v])
        t0 = time.perf_counter()
        oracle_graph.optimize(probe["max_iterations"])
        dt = time.perf_counter() - t0
        diff = 0.0
        for v in range(probe["n_vertices"]):
            e = oracle_graph.get_se3(v) if self.kinds[v] == 0 else oracle_graph.get_point_xyz(v)
            diff = max(diff, float(np.abs(np.asarray(e) - np.asarray(probe["after"][v])).max()))
        return 1e3 * dt, diff, int(oracle_graph.iterations)


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement of g2o (sparse-Cholesky LM) and PCL (RANSAC) on the host cores."""
    if rank != 0:
        return
    import oracle
    from semantic_slam_b200 import synth
    threads = os.cpu_count() or 1
    spec = synth.make_config_graph("cfg2")
    # same configuration as the GPU arm: the full 20 LM iterations per step.  Bounded sample = fewer STEPS when
    # steps x ~5.5 s would not end within a few minutes (the metric is a rate; every step does identical work)
    n_it = LM_ITERS
    warm_ref = min(args.warmup, 1)
    steps_ref = max(1, min(args.steps, int(200.0 / 5.5) - warm_ref))
    total = steps_ref + warm_ref
    args = argparse.Namespace(**{**vars(args), "warmup": warm_ref})
    times = []
    its = 0
    for s in range(total):
        o = oracle.OracleGraphSLAM(threads=threads)
        synth.load_graph(o, spec)
        t0 = time.perf_counter()
        o.optimize(n_it)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
            its += o.iterations
    T = sum(times)
    value = its / T
    # RANSAC sample: 8 crops x 1024 hypotheses of the cfg3 frame
    cl = synth.make_cloud()
    nbs = 8
    t0 = time.perf_counter()
    oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes[:nbs],
                              cl.triples[:nbs], want_counts=False, want_mask=False)
    tr = time.perf_counter() - t0
    npts = int((cl.boxes[:nbs, 2] * cl.boxes[:nbs, 3]).sum())
    sample = (f"cfg2 graph, all {n_it} LM iterations per step incl. symbolic analysis (g2o redoes it per optimize()), "
              f"{steps_ref} timed step(s) after {warm_ref} warm-up; edge linearisation on {threads} threads, sparse Cholesky "
              f"single-threaded (CSparse is)")
    line = {
        "impl": "reference", "metric": "LM iters/sec (10k-KF graph)", "value": value, "unit": "LM iters/s",
        "n_gpus": args.gpus, "steps": steps_ref, "warmup": warm_ref, "ms_per_step": 1e3 * T / max(len(times), 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: 10000 KF / %d landmarks / %d edges, %d LM iterations per step (synthetic, "
                               "lawn-mower trajectory, seed 20260927)" % (spec.n_landmarks, spec.n_edges, LM_ITERS),
                   "lm_iterations": n_it, "solver": "sparse Cholesky (CSparse restatement), as g2o lm_var"},
        "cpu_baseline": {"value": value, "unit": "LM iters/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "LM iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ransac": {"metric": "RANSAC Mpts/sec", "value": npts / tr / 1e6, "unit": "Mpts/s", "cores": 1,
                   "sample": f"{nbs} of 64 crops x 1024 hypotheses, PCL-order float arithmetic, 1 thread"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pcg-tol", type=float, default=1e-6)
    ap.add_argument("--preconditioner", type=int, default=3)
    ap.add_argument("--no-cfg4", action="store_true", help="skip the 100k-keyframe secondary workload (BASELINE.json configs[3])")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the per-frame loop (BASELINE.json configs[4])")
    ap.add_argument("--cfg5-frames", type=int, default=1000)
    ap.add_argument("--no-marginals", action="store_true", help="skip the landmark-marginals (K5) sub-object")
    ap.add_argument("--no-cluster", action="store_true", help="skip the dormant clustering-chain sub-object")
    ap.add_argument("--marginals-sample", type=int, default=64, help="cfg2 landmarks whose marginals are also timed in the iterative form (one PCG solve per column)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from semantic_slam_b200 import GraphSLAM, PlaneSegmentation, CloudLayout, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA back-end has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---------------- workloads -----------------------------------------------------------------
    spec = synth.make_config_graph("cfg2")
    gb = graph_bytes(spec)
    g = GraphSLAM(device=local, pcg_tol=args.pcg_tol, preconditioner=args.preconditioner)
    ids_cfg2 = synth.load_graph(g, spec)
    P0, X0 = g.get_all(spec.n_poses, spec.n_landmarks)
    from semantic_slam_b200 import distributed as ssbd
    if world > 1:
        ssbd.attach(g)                     # ONE graph, sharded by keyframe range over the ranks
    g.snapshot()
    cl = synth.make_cloud()
    lay = CloudLayout(cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets)
    seg = PlaneSegmentation(device=local)
    npts = int((cl.boxes[:, 2].astype(np.int64) * cl.boxes[:, 3]).sum())   # whole frame (all ranks together)
    my_boxes = np.ascontiguousarray(cl.boxes[rank::world])                   # crops are independent: round-robin
    my_triples = np.ascontiguousarray(cl.triples[rank::world])
    seg.upload(cl.msg, lay, my_boxes, my_triples)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident leg (`value`) ------------------------------------------------------
    t_lm = t_pcg = 0.0
    pcg_its = trials = lm_its = launches = 0
    t_rs = t_cnt = 0.0
    clocks = ClockSampler(local)
    if rank == 0:                          # one sampler per job: rank 0 prints the line, its GPU is the one reported
        clocks.start()
    for s in range(warmup + steps):
        if s == warmup:
            sync_all()
            clocks.mark()
            l0 = seg.launch_count()
        g.restore()
        flush.zero_()                      # flush L2 between timed iterations (untimed)
        torch.cuda.synchronize()
        g.optimize_resident(LM_ITERS)      # timed on its own stream with CUDA events (ms_device)
        flush.zero_()
        torch.cuda.synchronize()
        seg.run_resident()
        ms_all, ms_cnt = seg.timing()
        if s >= warmup:
            st = g.stats
            t_lm += st["ms_device"] * 1e-3
            t_pcg += st["ms_pcg"] * 1e-3
            pcg_its += st["total_pcg_iters"]
            trials += st["total_trials"]
            lm_its += st["iterations"]
            launches += st["kernel_launches"]
            t_rs += ms_all * 1e-3
            t_cnt += ms_cnt * 1e-3
    sync_all()
    clk = clocks.stop()
    launches += seg.launch_count() - l0
    final_chi2 = g.stats["chi2_final"]

    # ---------------- end-to-end leg (host buffers through the C-ABI) -----------------------------
    # inputs live in page-locked host memory (as a driver handing frames to the backend would keep them)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    msg_h, boxes_h, triples_h = pin(cl.msg), pin(my_boxes), pin(my_triples)
    P0, X0 = pin(P0), pin(X0)
    e2e_steps = max(2, min(steps, 5))
    t_e2e = 0.0
    t_e2e_r = 0.0
    for s in range(1 + e2e_steps):
        g.set_all(P0, X0)
        g.invalidate()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.set_all(P0, X0)                  # host estimates in
        g.optimize(LM_ITERS)               # CSR build + H2D + LM + D2H of the estimates
        P1, X1 = g.get_all(spec.n_poses, spec.n_landmarks)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        res, counts, mask = seg.fit_planes(msg_h, lay, boxes_h, triples_h)
        dtr = time.perf_counter() - t0
        if s >= 1:
            t_e2e += dt
            t_e2e_r += dtr
    e2e_its = g.stats["iterations"] * e2e_steps
    e2e_trials = g.stats["total_trials"]

    # max over ranks
    if world > 1:
        tt = torch.tensor([t_lm, t_e2e, t_rs, t_e2e_r, t_pcg, t_cnt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_lm, t_e2e, t_rs, t_e2e_r, t_pcg, t_cnt = [float(x) for x in tt.tolist()]
        ll = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        launches = int(ll.item())

    hbm_peak, peak_src = load_peaks()
    value = lm_its / t_lm                  # one job: every rank reports the same iteration count
    achieved = pcg_its * gb["b_cg"] / t_pcg / 1e9 if t_pcg > 0 else 0.0
    # DRAM bytes of one launch of the dominant kernel: not measurable inside this run (it needs an ncu replay), so it is
    # quoted from the committed capture and labelled as such
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            tj = json.load(open(tp))
            traffic = tj.get("k_pcg_dram_bytes_per_launch")
            traffic_src = "not measured in this run: " + str(tj.get("source", "ncu capture under profiles/"))
        except Exception:
            traffic = None
    shard = [g.shard_info(world, r) for r in range(world)] if world > 1 else []
    ransac_bytes = 16 * npts + 20 * 1024 * len(cl.boxes)
    line = {
        "metric": "LM iters/sec (10k-KF graph)", "value": value, "unit": "LM iters/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_lm / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: 10000 KF / %d landmarks / %d edges, %d LM iterations per step (synthetic, "
                               "lawn-mower trajectory, seed 20260927)" % (spec.n_landmarks, spec.n_edges, LM_ITERS),
                   "parallelism": ("keyframe-range sharded: ONE graph over %d GPUs (own keyframes %s, ghosts %s per rank), boundary "
                                   "cells pushed into peer memory over NVLink, no collective call inside the LM loop; RANSAC crops "
                                   "round-robin" % (world, [i[1] - i[0] for i in shard], [i[2] - (i[1] - i[0]) for i in shard]))
                   if world > 1 else "single GPU",
                   "l2": "flushed between timed steps (256 MiB write)", "pcg_tol": args.pcg_tol,
                   "preconditioner": ["block-Jacobi", "block-Jacobi + per-CTA rigid-body coarse level", "block-Jacobi + 5-pose aggregates + per-CTA coarse level (3-level additive)", "block-Jacobi + 5-pose aggregates coupled exactly in two groups per CTA + per-CTA coarse level"][args.preconditioner],
                   "lm_iterations": lm_its // max(steps, 1), "trials_per_step": trials / max(steps, 1),
                   "pcg_iters_per_step": pcg_its / max(steps, 1), "chi2_final": final_chi2},
        "e2e": {"value": e2e_its / t_e2e, "unit": "LM iters/s", "h2d_bytes_per_step": gb["h2d"],
                "d2h_bytes_per_step": gb["d2h"] + 96 * (e2e_trials + 2), "ms_per_step": 1e3 * t_e2e / e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"kernel": "k_pcg_flow<148> (persistent data-flow Schur-complement PCG, coarse Gauss-Jordan included)" +
                               (", %d ranks" % world if world > 1 else ""), "bound": "hbm", "achieved": achieved,
                     "peak": hbm_peak * world, "unit": "GB/s", "frac": achieved / (hbm_peak * world), "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "note": "algorithmic bytes = pcg iterations x B_cg (%d B, SURVEY 8d) / CUDA-event time of the "
                             "kernel; the cfg2 working set (~25 MB) is L2-resident, so this is latency-, not "
                             "HBM-bound" % gb["b_cg"]},
        "ransac": {"metric": "RANSAC Mpts/sec (640x480, 64 crops x 1024 hypotheses)", "value": steps * npts / t_rs / 1e6,
                   "unit": "Mpts/s", "points_per_step": npts, "ms_per_step": 1e3 * t_rs / steps,
                   "e2e": {"value": e2e_steps * npts / t_e2e_r / 1e6, "unit": "Mpts/s",
                           "h2d_bytes_per_step": int(cl.msg.size + cl.triples.size * 4 + cl.boxes.size * 4),
                           "d2h_bytes_per_step": int(64 * 64 + 64 * 1024 * 4 + npts)},
                   "roofline": {"kernel": "k_count (point x hypothesis sweep)", "bound": "hbm",
                                "achieved": steps * ransac_bytes / t_cnt / 1e9 if t_cnt > 0 else 0.0, "peak": hbm_peak,
                                "unit": "GB/s", "frac": (steps * ransac_bytes / t_cnt / 1e9) / hbm_peak if t_cnt > 0 else 0.0,
                                "tests_per_s": steps * npts * 1024 / t_cnt if t_cnt > 0 else 0.0,
                                "note": "1024 hypotheses are reused per loaded point (~500 flop/B): the binding roof "
                                        "is FP32 issue rate, not HBM (SURVEY 8d)"}},
    }

    # ---------------- K5: GraphSLAM::computeLandmarkMarginals (graph_slam.cpp:221-234) ----------------------------------
    # the reference calls it after every optimise for ALL mapped landmarks (semantic_graph_slam.cpp:89,181-205).  Timed through
    # the C-ABI with host buffers, for all landmarks of cfg2 right after the e2e optimise above and for all landmarks of a
    # 1 000-keyframe graph (the per-frame loop's size), in the direct form (csrc/ssb_marg_direct.cuh: the odometry chain is
    # eliminated by a block-bidiagonal Cholesky, the dense 3 Nl x 3 Nl landmark system inverted by block Gauss-Jordan) and, on a
    # small sample, in the iterative form it replaces (one PCG solve per column, SSB_MARG_DIRECT=0) — each beside the oracle's
    # time for the same call (a CSparse-style factorisation of the full system + g2o's MarginalCovarianceCholesky recursion,
    # `method="g2o"`: the algorithm the reference's computeMarginals runs — not the three triangular solves per landmark the GPU
    # tests use as their checker, which would flatter the GPU).
    if world == 1 and not args.no_marginals:
        def timed_marginals(graph, vids, reps=3):
            graph.computeLandmarkMarginals(vids[:2])               # warm-up (buffers)
            best, M = None, None
            for _ in range(reps):
                t0 = time.perf_counter()
                M = graph.computeLandmarkMarginals(vids)
                dtm = time.perf_counter() - t0
                best = dtm if best is None else min(best, dtm)
            return best, M
        lm_all = ids_cfg2[spec.vkind == 1].astype(np.int32)
        os.environ["SSB_MARG_DIRECT"] = "1"
        t_mg, Mg = timed_marginals(g, lm_all)
        n_s = min(args.marginals_sample, lm_all.size)
        sample = lm_all[:: max(1, lm_all.size // n_s)][:n_s]
        os.environ["SSB_MARG_DIRECT"] = "0"
        t_it, Mit = timed_marginals(g, sample, reps=1)
        os.environ["SSB_MARG_DIRECT"] = "1"
        specm = synth.make_graph(1000, 100, seed=synth.SEED_BASE + 6)
        gm = GraphSLAM(device=local, pcg_tol=1e-8, preconditioner=args.preconditioner)
        idm = synth.load_graph(gm, specm)
        gm.optimize(10)
        lmm = idm[specm.vkind == 1].astype(np.int32)
        t_mm, Mm = timed_marginals(gm, lmm)
        os.environ["SSB_MARG_DIRECT"] = "0"
        t_mi, Mmi = timed_marginals(gm, lmm, reps=1)
        os.environ["SSB_MARG_DIRECT"] = "1"
        pos_all = {int(v): k for k, v in enumerate(lm_all)}
        sel_s = np.array([pos_all[int(v)] for v in sample])
        line["marginals"] = {
            "metric": "landmark marginals per second (3x3 blocks of H^-1, computeLandmarkMarginals)",
            "cfg2_all": {"landmarks": int(lm_all.size), "seconds": t_mg, "value": lm_all.size / t_mg, "unit": "landmarks/s",
                         "form": "direct (poses eliminated, dense %d x %d landmark system)" % (3 * lm_all.size, 3 * lm_all.size),
                         "iterative_form": {"landmarks": int(sample.size), "seconds": t_it, "value": sample.size / t_it,
                                            "unit": "landmarks/s", "pcg_solves": int(3 * sample.size),
                                            "max_rel_diff_vs_direct": float(np.abs(Mit - Mg[sel_s]).max() / np.abs(Mg).max())}},
            "kf1000_all": {"landmarks": int(lmm.size), "seconds": t_mm, "value": lmm.size / t_mm, "unit": "landmarks/s",
                           "form": "direct",
                           "iterative_form": {"landmarks": int(lmm.size), "seconds": t_mi, "value": lmm.size / t_mi, "unit": "landmarks/s",
                                              "columns_per_launch": "up to 16 (one CG recurrence per copy of the graph)",
                                              "max_rel_diff_vs_direct": float(np.abs(Mmi - Mm).max() / np.abs(Mm).max())}}}
        if rank == 0 and not args.no_cpu_baseline:
            import oracle
            om = oracle.OracleGraphSLAM(threads=1)
            synth.load_graph(om, specm)
            om.optimize(10)
            t0 = time.perf_counter()
            Mo = om.computeLandmarkMarginals(lmm, method="g2o")
            t_om = time.perf_counter() - t0
            line["marginals"]["kf1000_all"]["cpu_baseline"] = {"value": lmm.size / t_om, "unit": "landmarks/s", "cores": 1, "kind": "port",
                                                               "seconds": t_om, "sample": _MARG_CPU_NOTE}
            line["marginals"]["kf1000_all"]["max_rel_diff_vs_oracle"] = float(np.abs(Mm - Mo).max() / np.abs(Mo).max())
            line["marginals"]["_pending_cfg2_all"] = [int(v) for v in lm_all]
            line["marginals"]["_Mg"] = Mg
        del gm

    # ---------------- the dormant k-means -> ProjectInliers -> ConvexHull chain (plane_segmentation.cpp:261-477) ----------
    # cv::kmeans of a full-frame sized sample (300 000 x 3, K = 4, 10 attempts) and the whole chain on a synthetic crop,
    # through the C-ABI with host buffers, beside the sequential oracle.  Labels / centres are bit-identical (tests).
    if world == 1 and not args.no_cluster:
        from semantic_slam_b200 import PlaneClustering
        pcl = PlaneClustering(device=local)
        rs = np.random.RandomState(7)
        cs = rs.randn(4, 3)
        cs /= np.linalg.norm(cs, axis=1, keepdims=True)
        big = (cs[rs.randint(0, 4, 300000)] + 0.08 * rs.randn(300000, 3)).astype(np.float32)
        cloud_c, T_c = synth.make_cluster_scene()
        cloud_c = cloud_c.reshape(-1, 4)
        nrm_c = np.full((cloud_c.shape[0], 4), np.nan, dtype=np.float32)
        okc = np.isfinite(cloud_c[:, 2])
        nrm_c[okc, :3] = T_c[2, :3] + rs.randn(int(okc.sum()), 3).astype(np.float32) * 0.03
        pcl.computeKmeans(big, 4, rng_state=999)
        t0 = time.perf_counter()
        kg = pcl.computeKmeans(big, 4, rng_state=999)
        t_kg = time.perf_counter() - t0
        pcl.clusterAndSegmentAllPlanes(cloud_c, nrm_c, T_c)
        t0 = time.perf_counter()
        rg = pcl.clusterAndSegmentAllPlanes(cloud_c, nrm_c, T_c)
        t_cg = time.perf_counter() - t0
        line["clustering"] = {"metric": "dormant plane-clustering chain (cv::kmeans + ProjectInliers + ConvexHull)",
                              "kmeans_300k": {"samples": 300000, "K": 4, "attempts": 10, "ms": 1e3 * t_kg, "value": 0.3 / t_kg, "unit": "Msamples/s"},
                              "chain": {"points": int(cloud_c.shape[0]), "clusters": int(len(rg["clusters"])), "hull_rows": int(rg["rows"].shape[0]),
                                        "ms": 1e3 * t_cg}}
        if rank == 0 and not args.no_cpu_baseline:
            import oracle
            t0 = time.perf_counter()
            ko = oracle.kmeans(big, 4, rng_state=999)
            t_ko = time.perf_counter() - t0
            t0 = time.perf_counter()
            ro = oracle.cluster_planes(cloud_c, nrm_c, T_c)
            t_co = time.perf_counter() - t0
            line["clustering"]["kmeans_300k"]["cpu_baseline"] = {"ms": 1e3 * t_ko, "value": 0.3 / t_ko, "unit": "Msamples/s", "cores": 1, "kind": "port"}
            line["clustering"]["kmeans_300k"]["labels_and_centres_bit_identical"] = bool(
                np.array_equal(kg[1], ko[1]) and np.array_equal(kg[2].view(np.uint32), ko[2].view(np.uint32)))
            line["clustering"]["chain"]["cpu_baseline"] = {"ms": 1e3 * t_co, "cores": 1, "kind": "port"}
            line["clustering"]["chain"]["clusters_identical"] = bool(
                len(rg["clusters"]) == len(ro["clusters"]) and np.array_equal(rg["labels"], ro["labels"]) and
                np.array_equal(rg["clusters"]["n_points"], ro["clusters"]["n_points"]))
        del pcl

    # ---------------- cfg4 (BASELINE.json configs[3]): the 100k-keyframe graph, same sharding ---------------------
    # The first genuinely HBM-bound size (B_cg = 280.6 MB per PCG iteration): it does not fit on chip, so the streaming
    # PCG kernel runs (3 barriers per iteration, spanning all ranks when sharded).  5 LM iterations per step.
    if not args.no_cfg4:
        spec4 = synth.make_config_graph("cfg4")
        gb4 = graph_bytes(spec4)
        # pcg_tol 1e-10: the loosest decade that keeps this (much worse conditioned) graph within the 1e-5 parity bar
        # after any iteration count (tests/test_gpu_graph.py::test_cfg4_three_iterations_vs_oracle, scripts/cfg4_parity.py)
        g4 = GraphSLAM(device=local, pcg_tol=1e-10, preconditioner=args.preconditioner)
        synth.load_graph(g4, spec4)
        if world > 1:
            ssbd.attach(g4)
        g4.snapshot()
        it4, t4, tp4, pi4 = 0, 0.0, 0.0, 0
        n4 = 3
        for s in range(1 + n4):
            g4.restore()
            flush.zero_()
            torch.cuda.synchronize()
            g4.optimize_resident(5)
            if s >= 1:
                it4 += g4.stats["iterations"]
                t4 += g4.stats["ms_device"] * 1e-3
                tp4 += g4.stats["ms_pcg"] * 1e-3
                pi4 += g4.stats["total_pcg_iters"]
        if world > 1:
            tt = torch.tensor([t4, tp4], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t4, tp4 = [float(x) for x in tt.tolist()]
        ach4 = pi4 * gb4["b_cg"] / tp4 / 1e9 if tp4 > 0 else 0.0
        line["cfg4"] = {"metric": "LM iters/sec (100k-KF graph)", "value": it4 / t4, "unit": "LM iters/s", "scaling": "strong",
                        "workload": "cfg4: %d KF / %d landmarks / %d edges, 5 LM iterations per step, %d timed steps" %
                                    (spec4.n_poses, spec4.n_landmarks, spec4.n_edges, n4),
                        "pcg_tol": 1e-10, "ms_per_step": 1e3 * t4 / n4, "pcg_iters_per_step": pi4 / n4, "us_per_pcg_iteration": 1e6 * tp4 / max(pi4, 1),
                        "chi2_final": g4.stats["chi2_final"],
                        "roofline": {"kernel": "k_pcg (streaming Schur-complement PCG)" + (", %d ranks" % world if world > 1 else ""),
                                     "bound": "hbm", "achieved": ach4, "peak": hbm_peak * world, "unit": "GB/s",
                                     "frac": ach4 / (hbm_peak * world),
                                     "note": "algorithmic bytes = pcg iterations x B_cg (%d B) / CUDA-event time of the kernel" % gb4["b_cg"]}}
        del g4

    # ---------------- cfg5 (BASELINE.json configs[4]): the per-frame loop of semantic_graph_slam::run ------------------
    # (src/ps_graph_slam/semantic_graph_slam.cpp:58-102) every keyframe: segment the detections' crops (the LIVE path:
    # integral-image normals + organised multi-plane segmentation, ssb_organized_planes), associate, grow the graph,
    # re-optimise the WHOLE graph (optimize(1024) until g2o's LM terminates).  The stream's detections are camera-frame
    # centroids (synth.make_frame_stream), so the segmentation runs on a fixed synthetic frame with one crop per detection:
    # it is timed inside the loop, its output is checked against the oracle in tests/test_gpu_segment.py, and the
    # association consumes the stream's centroids.  Single GPU (a growing graph of <= 4000 keyframes fits one chip).
    if not args.no_cfg5 and world == 1:
        from semantic_slam_b200 import DataAssociation, OrganizedSegmentation, SemanticGraphSLAM
        n_kf = args.cfg5_frames
        stream = synth.make_frame_stream(n_kf, max(12, n_kf // 10), seed=synth.SEED_BASE + 5, max_det=3)
        clf = synth.make_cloud(n_boxes=3, n_hyp=1, nan_frac=0.0005, box_min=150, box_max=200, seed=77)
        layf = CloudLayout(clf.width, clf.height, clf.point_step, clf.row_step, clf.offsets)
        oseg = OrganizedSegmentation(device=local, num_point_seg=500)
        KITTI = dict(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1, strict=True)

        def drive(graph, assoc, n, segment):
            slam = SemanticGraphSLAM(graph, assoc, stream.info6, cam_angle=stream.cam_angle, max_iterations=1024)
            t_seg = t_all = 0.0
            marks = {}
            t0 = time.perf_counter()
            for k in range(n):
                if stream.detections[k]:
                    ts = time.perf_counter()
                    segment(len(stream.detections[k]))
                    t_seg += time.perf_counter() - ts
                slam.feed(stream.odom[k], stream.detections[k])
                slam.run()
                if k + 1 in (n // 4, n // 2, n, 250):
                    marks[k + 1] = time.perf_counter() - t0
            t_all = time.perf_counter() - t0
            return slam, t_all, t_seg, marks

        g5 = FrameRecorder(GraphSLAM(device=local, preconditioner=args.preconditioner, pcg_tol=args.pcg_tol),
                           [n_kf // 4, n_kf // 2, 3 * n_kf // 4, n_kf])
        slam5, t5, t5seg, marks5 = drive(g5, DataAssociation(**KITTI), n_kf,
                                         lambda nd: oseg.segment(clf.msg, layf, clf.boxes[:nd], max_regions=8))
        line["cfg5"] = {"metric": "frames/sec (per-frame segment + associate + optimise loop)", "value": n_kf / t5, "unit": "frames/s",
                        "workload": "cfg5: %d keyframes, %d mapped landmarks, %d edges at the end; optimize(1024) of the whole graph "
                                    "per keyframe" % (n_kf, len(slam5.landmark_nodes_), g5.num_edges()),
                        "ms_per_frame": 1e3 * t5 / n_kf, "segmentation_ms_per_frame": 1e3 * t5seg / n_kf,
                        "elapsed_s_at_frames": marks5, "pcg_tol": args.pcg_tol}
        if rank == 0 and not args.no_cpu_baseline:
            import oracle
            from oracle.association import OracleDataAssociation
            n_cpu = min(n_kf, 250)
            crops = [oracle.crop(clf.msg, clf.width, clf.height, clf.point_step, clf.row_step, clf.offsets, clf.boxes[b]) for b in range(3)]
            slamo, t5o, t5oseg, _ = drive(oracle.OracleGraphSLAM(threads=1), OracleDataAssociation(**KITTI), n_cpu,
                                          lambda nd: [oracle.organized_planes(crops[b], min_inliers=500) for b in range(nd)])
            same = slamo.association_log == slam5.association_log[:len(slamo.association_log)]
            line["cfg5"]["cpu_baseline"] = {"value": n_cpu / t5o, "unit": "frames/s", "cores": 1, "kind": "port",
                                            "sample": "the first %d keyframes of the same stream (%.1f s, of which segmentation %.1f s)"
                                                      % (n_cpu, t5o, t5oseg)}
            line["cfg5"]["gpu_on_the_same_%d_frames" % n_cpu] = {"value": n_cpu / marks5.get(n_cpu, t5 * n_cpu / n_kf) if n_cpu in marks5 else None,
                                                                  "unit": "frames/s"}
            line["cfg5"]["association_identical_to_cpu"] = bool(same)
            # the CPU loop above covers the cheap early frames only; the same frame's optimize() at four graph sizes:
            per_size = []
            for pr in g5.probes:
                cpu_ms, diff, its_o = g5.replay_on(oracle.OracleGraphSLAM(threads=1), pr)
                per_size.append({"keyframes": pr["frame"], "gpu_ms": pr["gpu_ms"], "cpu_ms": cpu_ms, "lm_iterations": pr["lm_iterations"],
                                 "lm_iterations_cpu": its_o, "trials": pr["trials"], "max_abs_diff_after": diff})
            line["cfg5"]["optimize_of_one_frame_by_graph_size"] = per_size
        del g5

    # ---------------- CPU baseline (rank 0, N == 1 only) -----------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        o = oracle.OracleGraphSLAM(threads=1)
        synth.load_graph(o, spec)
        t0 = time.perf_counter()
        o.optimize(LM_ITERS)
        dt = time.perf_counter() - t0
        Po, Xo = o.get_all(spec.n_poses, spec.n_landmarks)
        line["cpu_baseline"] = {"value": o.iterations / dt, "unit": "LM iters/s", "cores": 1, "kind": "port",
                                "sample": "cfg2 graph, full %d LM iterations, sparse-Cholesky LM restatement of g2o "
                                          "lm_var, 1 thread (%.1f s)" % (o.iterations, dt)}
        line["parity"] = {"max_abs_pose_diff_vs_oracle": float(np.abs(P1 - Po).max()),
                          "max_abs_landmark_diff_vs_oracle": float(np.abs(X1 - Xo).max()),
                          "oracle_chi2_final": float(o.history[-1, 1])}
        if "marginals" in line and "_pending_cfg2_all" in line["marginals"]:
            # K5 on cfg2: ALL landmarks on both sides, same end state (20 LM iterations on both sides)
            va = np.array(line["marginals"]["_pending_cfg2_all"], dtype=np.int32)
            t0 = time.perf_counter()
            Mo2 = o.computeLandmarkMarginals(va, method="g2o")
            dtm = time.perf_counter() - t0
            c2 = line["marginals"]["cfg2_all"]
            c2["cpu_baseline"] = {"value": va.size / dtm, "unit": "landmarks/s", "cores": 1, "kind": "port", "seconds": dtm,
                                  "sample": "ALL %d landmarks of cfg2; %s" % (va.size, _MARG_CPU_NOTE)}
            c2["max_rel_diff_vs_oracle"] = float(np.abs(line["marginals"]["_Mg"] - Mo2).max() / np.abs(Mo2).max())
        # the RANSAC half next to ITS CPU baseline (PCL-order restatement, 1 thread, 8 of the 64 crops)
        nbs = 8
        t0 = time.perf_counter()
        oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes[:nbs],
                                  cl.triples[:nbs], want_counts=False, want_mask=False)
        tr = time.perf_counter() - t0
        nps = int((cl.boxes[:nbs, 2].astype(np.int64) * cl.boxes[:nbs, 3]).sum())
        line["ransac"]["cpu_baseline"] = {"value": nps / tr / 1e6, "unit": "Mpts/s", "cores": 1, "kind": "port",
                                          "sample": "%d of 64 crops x 1024 hypotheses (%.1f s)" % (nbs, tr)}
        line["ransac"]["vs_cpu_baseline"] = {"resident": line["ransac"]["value"] / (nps / tr / 1e6),
                                             "e2e": line["ransac"]["e2e"]["value"] / (nps / tr / 1e6)}
    if "marginals" in line:
        line["marginals"].pop("_pending_cfg2_all", None)
        line["marginals"].pop("_Mg", None)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
