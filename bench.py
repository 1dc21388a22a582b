#!/usr/bin/env python
"""bench.py - collision-checked configurations / s and swept-edge checks / s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload all|configs|edges|blocks|clutter|sussman]

The JSON line's own fields are BASELINE.json configs[1] - one step = one pass of the validity check (FK +
broadphase + narrowphase + exact re-check of the contact band) over 2^20 random Baxter right-arm configurations per
GPU against the Sussman block scene.  With the default `--workload all` the other BASELINE.json configurations
ride in the same line under `secondary`, each with its own value / roofline / e2e / cpu_baseline:

    edges    configs[2]  100 K RRT-like swept edges per GPU at OMPL's 1 % resolution       (weak scaling)
    blocks   configs[3]  2^20 right-arm configurations, 32-block tabletop, mesh robot       (weak scaling)
    clutter  configs[4]  ONE batch of 2^24 dual-arm configurations vs 512 primitives, split over the N GPUs
                         in 1024-aligned slices (strong scaling: the batch is fixed, N = 1 runs all of it)
    fk       north_star kernel (1): batched forward kinematics alone, HBM roofline           (rank-local)
    sussman  configs[0]  the motion side of the Baxter Sussman task-motion plan, end to end (rank 0)

`value` is measured with the inputs resident in HBM through the asynchronous device-pointer entry points
(aa_rx_cl_check_batch_dev / _edges_dev; for N > 1 the sharded forms aa_rx_cl_check_*_shard_dev, whose
ncclAllGather of the validity words - inside libtmkcl.so - is part of every timed step); `e2e` goes through the
blocking host-buffer C call with the H2D / D2H copies inside the timed region.  torch is plumbing here: device
buffers, streams, and torch.distributed for the barrier, the max over ranks and for carrying the 128-byte
communicator id to the ranks.

The timed steps replay CUDA graphs the library builds for arguments it has seen before (the untimed warm-up goes
twice around the ring of input batches); per-kernel times in `roofline` come from a second, event-instrumented
loop.  `cpu_baseline` / `--impl reference` time the CPU restatement (oracle/, early-out mode) on the host cores -
the only use of oracle/ here.  It is the builder's own port: the reference's Amino / FCL / OMPL are not in the
reference tree, so parity with them is UNPINNED; `amino_fcl` in the line reports the probe for a real
installation on the box.
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

# stdout carries the one JSON line and nothing else: library banners written to fd 1 (NCCL prints its version
# there) go to stderr instead
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CFG = 1 << 20
N_EDGE = 100_000
N_CLUTTER = 1 << 24
CHUNK = 1 << 20  # the strong-scaled batch is generated chunk by chunk so that its content does not depend on N
SECONDARY = ("edges", "blocks", "clutter", "fk", "sussman")
PARITY_NOTE = ("CPU arm = the builder's restatement of Amino/FCL/OMPL semantics (oracle/); the reference's own "
               "implementation of this path is not in the reference tree: parity with it is unpinned")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                          ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [x.strip() for x in line.split(",")]
            self.sm.append(float(r[0])); self.mx.append(float(r[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.002 if self.nvml else 0.1)

    def sample_now(self):
        """one sample from the calling thread (the timed loop is short and keeps the interpreter busy)"""
        try:
            self._sample_nvml() if self.nvml else self._sample_smi()
        except Exception:
            pass

    def start(self):
        self.t.start()

    def stop(self):
        self._stop.set()
        self.t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def flops_per_config(st, n_joint_rev):
    """F = 61 N_mov + 34 N_rev + 33 G_mov + 11 P + sum_t C_t n_t (n_t measured)."""
    n = max(1, st.n_states)
    C = (54.0, 300.0, 1000.0, 250.0)  # closed forms (mean of 48/60), box-box SAT, GJK pair, mesh leaf
    narrow = sum(C[k] * st.n_narrow[k] for k in range(4)) / n
    return 61.0 * st.n_joints + 34.0 * n_joint_rev + 33.0 * st.n_moving_geoms + 11.0 * st.n_pairs + narrow


# workload -> (scene builder, varied joints, units, kind, scaling, ring, BASELINE.json config)
def workload_table():
    from tmkit_b200 import scenes as S
    return {
        "configs": (lambda: S.baxter_sussman(), S.RIGHT_ARM, N_CFG, "configs", "weak", 8,
                    "BASELINE configs[1]: 2^20 random Baxter 7-DOF right-arm configurations vs the Sussman block scene "
                    "(restated Baxter fixture), per GPU"),
        "edges": (lambda: S.baxter_sussman(), S.RIGHT_ARM, N_EDGE, "edges", "weak", 8,
                  "BASELINE configs[2]: 100K random RRT-like Baxter right-arm segments, OMPL 1% resolution, per GPU"),
        "blocks": (lambda: S.blocksworld_scene(32, seed=32), S.RIGHT_ARM, N_CFG, "configs", "weak", 8,
                   "BASELINE configs[3]: 2^20 right-arm configurations, 32-block tabletop, one block in the gripper, "
                   "mesh robot vs box obstacles, per GPU"),
        "clutter": (lambda: S.clutter_scene(512, seed=0), S.RIGHT_ARM + S.LEFT_ARM, N_CLUTTER, "configs", "strong", 2,
                    "BASELINE configs[4]: ONE batch of 2^24 dual-arm (14-DOF) configurations vs 512 random primitives "
                    "(obstacles touching the robot at q0 rejected), split over the GPUs in 1024-aligned slices"),
    }


def metric_name(workload):
    return {"configs": "collision-checked configs/sec (FK + collision, Baxter right arm vs Sussman scene)",
            "edges": "swept-edge checks/sec (RRT-Connect segments at OMPL resolution)",
            "blocks": "collision-checked configs/sec (FK + collision, Baxter right arm, 32-block tabletop)",
            "clutter": "collision-checked configs/sec (FK + collision, dual-arm Baxter vs 512 primitives)",
            "sussman": "Baxter Sussman motion-plan time (6 motion queries: IK + RRT-Connect + simplify + reparent)"}[workload]


def config_dict(workload):
    """identical in both arms (the driver compares them)"""
    _, _, n_units, kind, scaling, ring, desc = workload_table()[workload]
    key = "units_per_step_per_gpu" if scaling == "weak" else "units_per_step_total"
    l2 = (f"inputs rotate over a ring of {ring} distinct batches"
          + (" (> L2)" if kind == "configs" else " (11 MB each: L2-resident, as a planner's frontier is)"))
    return {"workload": desc, key: n_units, "l2": l2}


def make_problem(workload):
    from tmkit_b200 import amino as A
    from tmkit_b200 import scenes as S
    from tmkit_b200 import workloads as W
    build, names, n_units, kind, scaling, ring, desc = workload_table()[workload]
    return build(), S, W, A, names, n_units, kind, scaling, ring


def strong_slice(W, lim, lo, hi, ring_index):
    """rows lo .. hi-1 of ring batch `ring_index` of the strong-scaled workload; chunk c of the batch is always
    random_configs(seed = 64 * ring_index + c), whatever N is"""
    out = []
    c = lo // CHUNK
    while c * CHUNK < hi:
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        Q = W.random_configs(lim, CHUNK, seed=5000 + 64 * ring_index + c)
        out.append(Q[a - c * CHUNK:b - c * CHUNK])
        c += 1
    return np.concatenate(out) if out else np.zeros((0, len(lim)))


# ------------------------------------------------------------------------------------------------ CPU arm
_ORACLE_SO = None


def oracle_lib():
    """the CPU restatement, built for this box's cores into oracle/ (in-tree, so that the driver sees what the
    reference arm loaded)"""
    global _ORACLE_SO
    from oracle import oracle as O
    if _ORACLE_SO is None:
        try:
            _ORACLE_SO = O.build(force=True, march_native=True, out=os.path.join(ROOT, "oracle", "libtmkoracle_native.so"))
        except Exception:
            _ORACLE_SO = O.build()
    return O, _ORACLE_SO


def cpu_problem(workload):
    O, so = oracle_lib()
    sc, S, W, _, names, n_units, kind, scaling, ring = make_problem(workload)
    orc = O.OracleScene(S.flatten(sc), libpath=so)
    q0 = np.asarray([S.Q0.get(n, 0.0) for n in orc.config_names])
    orc.allow_config(q0)
    sub = orc.config_ids(names)
    lim = W.arm_limit_table(names)
    cores = os.cpu_count() or 1
    edges = kind == "edges"

    def make(n, seed):
        return W.random_edges(lim, n, seed=seed) if edges else (W.random_configs(lim, n, seed=seed),)

    def run(batch):
        t0 = time.perf_counter()
        if edges:
            orc.check_edges(sub, q0, batch[0], batch[1], W.seg_len(lim), early_out=True, threads=cores)
        else:
            orc.check_batch(sub, q0, batch[0], early_out=True, threads=cores)
        return time.perf_counter() - t0

    return make, run, cores, edges, n_units


def cpu_baseline(workload: str, seconds_budget: float = 10.0):
    """The CPU restatement (early-out mode = aa_rx_cl_check with a NULL set) on all host cores, on a bounded
    sample of the same workload."""
    make, run, cores, edges, n_units = cpu_problem(workload)
    n_cal = 2000 if edges else 4000
    rate = n_cal / max(run(make(n_cal, 0)), 1e-6)
    n2 = int(min(n_units, 200_000 if edges else 4_000_000, max(n_cal, rate * seconds_budget)))
    dt = run(make(n2, 1))
    what = "RRT-like edges" if edges else "configurations"
    return {"value": n2 / dt, "unit": "edges/s" if edges else "configs/s", "cores": cores, "kind": "port",
            "sample": f"{n2} of {n_units} {what}, early-out, bounding-sphere culls, all cores",
            "note": "CPU restatement of Amino/FCL semantics (oracle, early-out mode); not Amino/FCL itself"}


def amino_probe():
    """is a real Amino / FCL installation on this box?  If so oracle/_ref (make -C oracle ref) builds
    the check_state loop of the reference's demo/tmpexec.c:362-428 against it; otherwise parity stays unpinned."""
    try:
        have = {p: subprocess.run(["pkg-config", "--exists", p]).returncode == 0 for p in ("amino", "amino-collision", "fcl", "ompl")}
    except Exception as e:
        return {"present": False, "probe": f"pkg-config unavailable: {e}"}
    out = {"present": bool(have["amino"] and have["fcl"]), "probe": "pkg-config --exists " + " ".join(have), "found": have}
    if out["present"]:
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], capture_output=True, text=True)
        out["ref_build"] = "ok" if r.returncode == 0 else (r.stdout + r.stderr)[-400:]
    return out


def reference_line(workload, steps, warmup, gpus, budget_s):
    """K timed steps after W warm-up steps of the CPU restatement, each step a bounded sample of the workload."""
    make, run, cores, edges, n_units = cpu_problem(workload)
    unit = "edges/s" if edges else "configs/s"
    n_cal = 2000 if edges else 4000
    rate = n_cal / max(run(make(n_cal, 0)), 1e-6)
    per_step = max(0.05, budget_s / max(1, steps + warmup))
    if os.environ.get("TMK_BENCH_REF_BUDGET"):  # test aid
        per_step = float(os.environ["TMK_BENCH_REF_BUDGET"])
    n_step = int(max(n_cal, min(n_units if not edges else N_EDGE, rate * per_step)))
    if workload == "clutter":
        n_step = int(max(n_cal, min(N_CFG, rate * per_step)))
    ring = [make(n_step, 1 + r) for r in range(2)]
    for w in range(warmup):
        run(ring[w % 2])
    total = float(sum(run(ring[i % 2]) for i in range(steps)))
    v = n_step * steps / total
    what = "RRT-like edges" if edges else "configurations"
    whole = "the whole batch" if n_step == n_units else f"{n_step} of {n_units}"
    sample = f"{whole} {what} per step, early-out, bounding-sphere culls, all cores"
    _, _, _, _, scaling, _, _ = workload_table()[workload]
    return {
        "impl": "reference", "sample_units": n_step, "sample_s": total / steps, "metric": metric_name(workload),
        "value": v, "unit": unit, "n_gpus": gpus, "steps": steps, "warmup": warmup, "ms_per_step": total / steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(workload),
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "parity": PARITY_NOTE,
    }


def run_reference(args):
    """The reference arm: the CPU restatement on all host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    primary = "configs" if args.workload == "all" else args.workload
    if primary == "sussman":
        line = run_sussman(args, cpu_only=True)
        line["impl"] = "reference"
        emit(line)
        return
    line = reference_line(primary, args.steps, args.warmup, args.gpus, 120.0)
    if args.workload == "all":
        sec = {}
        for wl in SECONDARY:
            if wl in ("sussman", "fk"):
                continue  # sussman needs the GPU planner's edge log (its CPU replay rides in the own arm); fk has no CPU arm
            sec[wl] = reference_line(wl, 3, 1, args.gpus, 12.0)
        line["secondary"] = sec
    line["amino_fcl"] = amino_probe()
    emit(line)


# ------------------------------------------------------------------------------------------------ C1
def run_sussman(args, steps=None, cpu_only=False):
    """BASELINE configs[0]: the motion side of the Baxter Sussman task-motion plan, end to end
    (tmkit_b200/sussman.py), beside the CPU restatement replaying exactly the same validity work."""
    import torch
    from tmkit_b200 import scenes as S
    from tmkit_b200 import sussman as SU
    from tmkit_b200 import workloads as W
    O, so = oracle_lib()
    if not torch.cuda.is_available():
        if cpu_only:
            return {"metric": metric_name("sussman"), "unavailable": "the edge log of the plan comes from the GPU planner"}
        raise SystemExit("bench.py needs a CUDA device: the validity check has no CPU fallback")
    steps = steps or args.steps
    for w in range(max(1, min(args.warmup, 3))):
        SU.run(seed=900 + w)
    runs = []
    for k in range(steps):
        log = []
        r = SU.run(seed=k, edge_log=log)
        runs.append((r, log))
    ok = [r for r, _ in runs if r.ok]
    tot = float(np.median([r.times["total"] for r in ok])) if ok else None
    # CPU baseline: the oracle (early-out) on the edges of one run, one thread (the reference's planner is
    # single-threaded) and all cores
    r, log = next(((r, log) for r, log in runs if r.ok), runs[0])
    cores = os.cpu_count() or 1
    cpu = {}
    n_edges = 0
    for label, th in (("1", 1), ("all", cores)):
        t = 0.0
        for desc, q_start, names, q0e, q1e in log:
            orc = O.OracleScene(S.flatten(desc), libpath=so)
            orc.allow_config(q_start)
            lim = W.arm_limit_table(names)
            t0 = time.perf_counter()
            orc.check_edges(orc.config_ids(names), q_start, q0e, q1e, W.seg_len(lim), early_out=True, threads=th)
            t += time.perf_counter() - t0
            if label == "1":
                n_edges += len(q0e)
        cpu[label] = t
    split = {k: float(np.median([x.times[k] for x in ok])) for k in ("ik", "rrt", "simplify", "scene")} if ok else {}
    return {
        "metric": metric_name("sussman"),
        "value": tot, "unit": "s", "n_gpus": 1, "steps": steps, "warmup": args.warmup,
        "ms_per_step": None if tot is None else tot * 1e3, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "restated Baxter fixture + reference Sussman scene files",
        "config": {"workload": "BASELINE configs[0]: demo/baxter-sussman, task plan fixed (task planner out of scope), "
                               "motion timeout 3 s, simplify on", "success": f"{len(ok)}/{len(runs)}"},
        "split_s": split,
        "counters": r.counters,
        "cpu_baseline": {"value": cpu["all"], "unit": "s", "cores": cores, "kind": "port",
                         "sample": f"oracle replay of the {n_edges} edges one plan checked (validity work only)",
                         "single_thread_s": cpu["1"]},
        "validity_s_gpu": split.get("rrt", 0.0) + split.get("simplify", 0.0),
        # the whole plan already runs through the host-buffer C calls (aa_rx_mp_*): value IS end to end
        "e2e": {"value": tot, "unit": "s", "h2d_bytes_per_step": int(2 * 8 * 7 * r.counters["edges_checked"]),
                "d2h_bytes_per_step": int(4 * ((r.counters["edges_checked"] + 31) // 32 + r.counters["edge_calls"])),
                "call": "aa_rx_mp_set_start / _set_wsgoal / _plan per action", "note": "bytes: edge end points in, validity words out"},
    }


# ------------------------------------------------------------------------------------------------ K1 alone
def run_fk(ctx, args, steps):
    """aa_rx_sg_tf_batch_dev (north_star kernel 1, SURVEY.md 8a row a5 batched): 2^20 right-arm configurations, world
    poses of the 17 frames that move with the arm, float planes out.  HBM-bound: 28 B in + 17 x 28 B out per unit."""
    torch = ctx.torch
    from tmkit_b200 import amino as A
    from tmkit_b200 import scenes as S
    from tmkit_b200 import workloads as W
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q0 = S.q0_vector(sc)
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    n = N_CFG
    # the frames the varied joints move: everything below right_s0 in the tree
    names = sg.frame_names()
    moving, frames = {sg.frame_id("right_s0")}, []
    for i in range(sg.frame_count()):
        if i in moving or sg.frame_parent(i) in moving:
            moving.add(i)
            frames.append(i)
    ring = [torch.from_numpy(np.ascontiguousarray(W.random_configs(lim, n, seed=300 + r).T.astype(np.float32))).to(ctx.dev) for r in range(4)]
    out = torch.zeros((len(frames) * 7, n), dtype=torch.float32, device=ctx.dev)
    rows = torch.zeros((n, len(frames), 7), dtype=torch.float64, device=ctx.dev)
    ts = torch.cuda.Stream(device=ctx.dev)

    def timed(fn, k):
        with torch.cuda.stream(ts):
            for i in range(3):
                fn(i)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ts.synchronize()
            ev[0].record(ts)
            for i in range(k):
                fn(i)
            ev[1].record(ts)
            ts.synchronize()
        return ev[0].elapsed_time(ev[1]) / k

    ms = timed(lambda i: sg.tf_batch_dev(n, sub, q0, ring[i % 4].data_ptr(), A.Q_SOA_F32, n, frames, out.data_ptr(), A.TF_SOA_F32, n,
                                         ts.cuda_stream), steps)
    ms_rows = timed(lambda i: sg.tf_batch_dev(n, sub, q0, ring[i % 4].data_ptr(), A.Q_SOA_F32, n, frames, rows.data_ptr(), A.TF_ROWS_F64, 0,
                                              ts.cuda_stream), max(3, steps // 4))
    # end to end: host double rows in, host double transforms out (aa_rx_sg_tf_batch), a 2^16 sample
    m = 1 << 16
    Qh = W.random_configs(lim, m, seed=9)
    sg.tf_batch(sub, q0, Qh[:1024], frames=frames)
    t0 = time.perf_counter()
    sg.tf_batch(sub, q0, Qh, frames=frames)
    e2e = m / (time.perf_counter() - t0)
    ctx.barrier()
    if ctx.rank != 0:
        return None
    hbm_peak, peak_kind, _ = load_peaks()
    bytes_unit = 4.0 * len(sub) + 28.0 * len(frames)
    gbs = bytes_unit * n / (ms * 1e-3) / 1e9
    flops_unit = 61.0 * len(frames) + 34.0 * len(sub)  # SURVEY.md 8d FK term: 61 per moving frame + 34 per revolute joint
    return {
        "metric": "batched forward kinematics, configs/sec (aa_rx_sg_tf_batch_dev, 17 moving frames of the right arm)",
        "value": n / (ms * 1e-3), "unit": "configs/s", "n_gpus": 1, "steps": steps, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "dtype": "f32", "data": "synthetic",
        "config": {"workload": "north_star kernel (1): 2^20 right-arm configurations, SoA float joints in, 17 x 7 float planes out",
                   "frames": len(frames)},
        "value_f64_rows_out": n / (ms_rows * 1e-3),
        "e2e": {"value": e2e, "unit": "configs/s", "h2d_bytes_per_step": m * len(sub) * 8, "d2h_bytes_per_step": m * len(frames) * 56,
                "call": "aa_rx_sg_tf_batch (host double rows in and out, 2^16 configurations)"},
        "gpu_launches": steps,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                     "kernel": "k_fk_frames<TMK_TF_SOA_F32>", "kernel_ms": ms, "bytes_per_unit": bytes_unit,
                     "algorithmic_bytes": bytes_unit * n, "peak_kind": peak_kind, "flops_per_unit": flops_unit},
    }


# ------------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """process-wide plumbing: device, torch.distributed (barrier / max over ranks / id broadcast), the library's
    shard communicator"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the validity check has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.comm = None
        if self.world > 1:
            from tmkit_b200 import amino as A
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            idt = torch.zeros(A.SHARD_ID_BYTES, dtype=torch.uint8, device=self.dev)
            if self.rank == 0:
                idt.copy_(torch.frombuffer(bytearray(A.ShardComm.unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, src=0)
            self.comm = A.ShardComm(self.rank, self.world, bytes(idt.cpu().numpy().tobytes()))
            # the gather of step k on the communicator's own high-priority stream, beside the kernels of step k+1
            self.overlap = os.environ.get("TMK_BENCH_GATHER", "overlap") == "overlap"
            self.comm.overlap(self.overlap)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.comm.destroy()
            self.dist.destroy_process_group()


def run_workload(ctx: Ctx, args, workload: str, steps: int, warmup: int, cpu_budget: float):
    torch, dist = ctx.torch, ctx.dist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    sc, S, W, A, names, n_units, kind, scaling, RING = make_problem(workload)
    sg = A.SceneGraph.from_scene(sc)
    q0 = np.asarray([S.Q0.get(n, 0.0) for n in sg.config_names()])
    sg.allow_config(q0)  # the reference allows what collides at the start configuration
    cl = A.Collision(sg)
    sub = sg.config_indices(names)
    lim = W.arm_limit_table(names)
    seg = W.seg_len(lim)
    edges = kind == "edges"
    unit = "edges/s" if edges else "configs/s"

    # ---- the batch and this rank's slice of it ----
    if scaling == "strong":
        n_total = n_units
    else:
        # weak: the batch grows with N.  Slices are 1024-multiples, so for the 100 K-edge workload every rank but
        # the last checks 100 352 edges (units_per_rank in the line); 2^20 configurations per rank are exact.
        n_total = A.lib().tmk_shard_per_rank(n_units, 1) * (world - 1) + n_units
    lo, hi = A.shard_bounds(n_total, world, rank)
    n_mine = hi - lo
    wpr = A.shard_words(n_total, world)
    nw_all = wpr * world if world > 1 else (n_total + 31) // 32

    # ---- inputs: a ring of distinct batches, SoA float32 in HBM (device-resident leg) ----
    ring, ring1, host_rows, host_rows1 = [], [], [], []
    for r in range(RING):
        if edges:
            a, b = W.random_edges(lim, n_mine, seed=1000 * rank + r)
            ring.append(torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).to(dev))
            ring1.append(torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).to(dev))
            if r < 2:
                host_rows.append(torch.from_numpy(a).pin_memory()); host_rows1.append(torch.from_numpy(b).pin_memory())
        else:
            Q = strong_slice(W, lim, lo, hi, r) if scaling == "strong" else W.random_configs(lim, n_mine, seed=1000 * rank + r)
            ring.append(torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).to(dev))
            if r < (1 if scaling == "strong" else 2):
                host_rows.append(torch.from_numpy(Q).pin_memory())
            del Q
    # two result buffers: the gather of step k (communicator stream) overlaps the kernels of step k+1
    bits2 = [torch.zeros(nw_all, dtype=torch.int32, device=dev) for _ in range(2)]
    ts = torch.cuda.Stream(device=dev)
    stream = ts.cuda_stream
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    sub_c = np.ascontiguousarray(sub, dtype=np.int32)
    q0_c = np.ascontiguousarray(q0, dtype=np.float64)
    sub_ptr, q0_ptr, n_sub_c, n_q_c, seg_c = sub_c.ctypes.data, q0_c.ctypes.data, len(sub_c), len(q0_c), float(seg)
    ring_ptr = [t.data_ptr() for t in ring]
    ring1_ptr = [t.data_ptr() for t in ring1]
    bits_ptr = [t.data_ptr() for t in bits2]
    L = cl.L
    comm_h = ctx.comm.h if ctx.comm else None

    def step(i):
        k = i % RING
        b = i & 1 if world > 1 else 0
        # the C entry points directly, arguments marshalled once: the loop below has to keep a ~1 ms GPU step fed
        if world == 1:
            if edges:
                L.aa_rx_cl_check_edges_dev(cl.h, n_mine, n_sub_c, sub_ptr, n_q_c, q0_ptr, ring_ptr[k], ring1_ptr[k],
                                           A.Q_SOA_F32, n_mine, seg_c, bits_ptr[b], None, stream)
            else:
                L.aa_rx_cl_check_batch_dev(cl.h, n_mine, n_sub_c, sub_ptr, n_q_c, q0_ptr, ring_ptr[k], A.Q_SOA_F32, n_mine,
                                           bits_ptr[b], stream)
            return
        # sharded form: this rank's slice into slot `rank` of the gathered bitmap + the library's ncclAllGather.
        # In overlap mode the gather of step i runs on the communicator's stream beside the kernels of step i+1
        # (other bitmap); the library makes step i+2 wait for it before bits2[b] is written again.
        if edges:
            L.aa_rx_cl_check_edges_shard_dev(cl.h, comm_h, n_total, n_sub_c, sub_ptr, n_q_c, q0_ptr, ring_ptr[k], ring1_ptr[k],
                                             A.Q_SOA_F32, n_mine, seg_c, bits_ptr[b], None, stream)
        else:
            L.aa_rx_cl_check_batch_shard_dev(cl.h, comm_h, n_total, n_sub_c, sub_ptr, n_q_c, q0_ptr, ring_ptr[k], A.Q_SOA_F32,
                                             n_mine, bits_ptr[b], stream)

    def drain():
        """the timed stream waits for the gathers still in flight: they belong to the timed steps"""
        if world > 1 and ctx.overlap:
            L.tmk_shard_comm_join(comm_h, stream)

    with torch.cuda.stream(ts):
        l2_flush.zero_()
        # the device-pointer entry points capture a CUDA graph the second time they see the same arguments: the
        # untimed warm-up goes twice around the ring so that the timed steps replay graphs (steady state)
        n_warm = max(warmup, 2 * RING)
        n_warm += n_warm & 1  # keeps the (ring slot, result buffer) pairing of the warm-up in the timed steps
        for i in range(n_warm):
            step(i)
        drain()
        ctx.barrier()
        sampler = ClockSampler(ctx.local)
        if rank == 0:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        cl.kernel_timing(False)
        ctx.barrier()
        ev[0].record(ts)
        t_host = time.perf_counter()
        for i in range(steps):
            step(n_warm + i)
        drain()
        ev[1].record(ts)
        host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / steps
        if rank == 0:
            # graph replays put the host far ahead of the GPU: the clocks are sampled now, with everything
            # enqueued and the GPU still working through the timed steps (an NVML query inside the loop would
            # stall the enqueueing thread for milliseconds)
            sampler.sample_now()
        ctx.barrier()
        ms_total = ev[0].elapsed_time(ev[1])
        # events around the dominant kernels, recorded by the library on `stream`.  The timed steps are replayed
        # from CUDA graphs (no events inside a graph): the kernel times come from this second, untimed loop
        k_steps = min(steps, 20)
        cl.kernel_timing(True)
        for i in range(k_steps):
            step(n_warm + i)
        drain()
        ctx.barrier()
        n_timed, k_ms_sum, rk_ms_sum = cl.kernel_time()
        cl.kernel_timing(False)
        clocks = sampler.stop() if rank == 0 else None
    assert n_timed == k_steps, (n_timed, k_steps)
    k_ms = k_ms_sum / n_timed
    rk_ms = rk_ms_sum / n_timed

    per_rank = None
    if world > 1:
        # per-rank step time, kernel time and host enqueue time (diagnostics beside the max the contract asks for)
        mine = torch.tensor([ms_total / steps, k_ms + rk_ms, host_enqueue_ms], dtype=torch.float64, device=dev)
        allr = torch.zeros(3 * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = [[round(float(x), 4) for x in allr[3 * r:3 * r + 3].tolist()] for r in range(world)]
    ms_step = ctx.max_over_ranks(ms_total) / steps
    units_step = n_total
    value = units_step / (ms_step * 1e-3)
    words = bits2[(n_warm + steps - 1) & 1 if world > 1 else 0].cpu().numpy().view(np.uint32)
    mine_words = words[rank * wpr:rank * wpr + (n_mine + 31) // 32] if world > 1 else words
    valid_frac = float(np.unpackbits(mine_words.view(np.uint8), bitorder="little")[:n_mine].mean()) if n_mine else None

    # ---- e2e: host buffers through the public C-ABI call (H2D + kernels + D2H inside), this rank's slice ----
    e2e_steps = max(3, min(steps, 10)) if scaling == "weak" else 3
    nw_mine = (n_mine + 31) // 32
    hb = np.zeros(nw_mine + 1, dtype=np.uint32)
    n_sub_i, n_q_i = len(sub), len(q0)

    def host_call(rows, rows1=None, f32=False):
        p = rows.data_ptr() if hasattr(rows, "data_ptr") else rows.ctypes.data
        if edges:
            p1 = rows1.data_ptr() if hasattr(rows1, "data_ptr") else rows1.ctypes.data
            L.aa_rx_cl_check_edges(cl.h, n_mine, n_sub_i, sub.ctypes.data, n_q_i, q0.ctypes.data, p, p1, n_sub_i, seg,
                                   hb.ctypes.data, None)
        elif f32:
            L.aa_rx_cl_check_batch_f32(cl.h, n_mine, n_sub_i, sub.ctypes.data, n_q_i, q0.ctypes.data, p, n_sub_i, hb.ctypes.data)
        else:
            L.aa_rx_cl_check_batch(cl.h, n_mine, n_sub_i, sub.ctypes.data, n_q_i, q0.ctypes.data, p, n_sub_i, hb.ctypes.data)

    def time_host(fn, n_steps):
        for i in range(2):
            fn(i)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(n_steps):
            fn(i)  # blocking: returns with the bitmap in host memory
        return units_step * n_steps / ctx.max_over_ranks(time.perf_counter() - t0)

    nh = len(host_rows)
    e2e_value = time_host(lambda i: host_call(host_rows[i % nh], host_rows1[i % nh] if edges else None), e2e_steps)
    h2d = n_mine * len(sub) * 8 * (2 if edges else 1)
    d2h = nw_mine * 4
    # beside it (information only): float rows - half the PCIe bytes - and pageable rows, what a caller that
    # malloc's its buffers (mp.c, a Lisp caller) hands over; aa_rx_host_alloc / _register pin them
    e2e_f32 = e2e_pageable = None
    if not edges:
        rows32 = [torch.from_numpy(np.ascontiguousarray(h.numpy().astype(np.float32))).pin_memory() for h in host_rows]
        e2e_f32 = time_host(lambda i: host_call(rows32[i % nh], f32=True), e2e_steps)
        del rows32
    if scaling == "weak":
        pageable = [np.array(h.numpy(), copy=True) for h in host_rows]
        pageable1 = [np.array(h.numpy(), copy=True) for h in host_rows1]
        e2e_pageable = time_host(lambda i: host_call(pageable[i % nh], pageable1[i % nh] if edges else None), e2e_steps)

    # ---- untimed runs for the counters: the production path (launches, states it evaluates, pairs it
    # re-checks in double) and the instrumented path (narrowphase calls by class for the flop formula) ----
    if edges:
        a, b = W.random_edges(lim, 20000, seed=7)
        cl.check_edges(sub, q0, a, b, seg, want_first=False)
    else:
        Qs = W.random_configs(lim, 1 << 16, seed=7)
        cl.check_batch(sub, q0, Qs)
    fast = cl.stats()
    launches_per_step = int(fast.n_launches)
    states_per_unit = fast.n_states / max(1, fast.n_units)
    unc_frac = fast.n_uncertain / max(1, fast.n_states)
    cl.stats_enable(True)
    if edges:
        cl.check_edges(sub, q0, a, b, seg, want_first=False)
    else:
        cl.check_batch(sub, q0, Qs)
    st = cl.stats()
    cl.stats_enable(False)
    n_rev = int(st.n_joints)
    f_cfg = flops_per_config(st, n_rev)
    states_per_unit_ref = st.n_states / max(1, st.n_units)
    fp32_probe = L.tmk_fp32_peak_tflops()  # measured on this device: dependent-free FFMA loop, all SMs
    del ring, ring1, host_rows, host_rows1, bits2, l2_flush
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    hbm_peak, peak_kind, sm_max = load_peaks()
    traffic = None  # dram bytes per launch of the dominant kernel, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(workload)
        if tj:
            traffic = tj["bytes_per_launch"] * (n_mine / tj["units_per_launch"])
    except Exception:
        traffic = None
    bytes_unit = (4 * len(sub) * (2 if edges else 1)) + 1.0 / 8
    achieved_gbs = bytes_unit * n_mine / (k_ms * 1e-3) / 1e9
    fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    flops_unit = f_cfg * states_per_unit
    achieved_tf = flops_unit * n_mine / (k_ms * 1e-3) / 1e12
    kernel = ("k_states_tile<SrcEdgeLevel> + k_heavy_mixed + k_edges_compact, all bisection levels" if edges
              else "wavefront pipeline k_wave_fk + k_wave_broad + k_wave_quick + k_wave_export + k_wave_bits, then the deferred "
                   "narrowphase (k_conv_items beside k_mesh_tasks -> k_tri_items)")
    par = f"dp{world}: contiguous 1024-aligned slice per rank"
    if world > 1:
        par += (", one ncclAllGather of the validity words per step inside libtmkcl.so (aa_rx_cl_check_*_shard_dev), "
                + ("on the communicator's stream, overlapping the next step; all gathers finish inside the timed region"
                   if ctx.overlap else "on the kernels' stream"))
    line = {
        "metric": metric_name(workload), "value": value, "unit": unit, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(workload), "parallelism": par,
        "units_per_rank": n_mine,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "aa_rx_cl_check_edges" if edges else "aa_rx_cl_check_batch", "host_dtype": "f64 rows, pinned",
                "value_f32_rows": e2e_f32, "value_pageable_f64_rows": e2e_pageable, "bytes": "per rank"},
        "gpu_launches": int(launches_per_step * steps),
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "per_rank_ms": per_rank,  # [step, kernels, host enqueue] per rank
        "clocks": clocks,
        # the path is bound by the FP32 (FMA) pipe, not HBM and not the tensor cores (SURVEY.md 8d,
        # DESIGN.md): `roofline` is the FP32 one, `roofline_hbm` the contract's HBM form beside it
        "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_probe, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp32_probe, "traffic": traffic, "kernel": kernel, "kernel_ms": k_ms,
                     "recheck_kernel_ms": rk_ms, "flops_per_unit": flops_unit, "flops_per_state": f_cfg,
                     "states_per_unit": states_per_unit,
                     "peak_kind": "measured on this device by tmk_fp32_peak_tflops (FFMA loop)",
                     "peak_nominal": fp32_nominal},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic, "algorithmic_bytes": bytes_unit * n_mine,
                         "peak_kind": peak_kind, "kernel": kernel,
                         "kernel_ms": k_ms, "bytes_per_unit": bytes_unit},
        "counters": {"joints": int(st.n_joints), "moving_geoms": int(st.n_moving_geoms), "pairs": int(st.n_pairs),
                     "narrow_per_state": [st.n_narrow[k] / max(1, st.n_states) for k in range(4)],
                     "states_per_unit_one_warp_per_edge_kernel": states_per_unit_ref,
                     "double_rechecked_pairs_per_state": unc_frac, "valid_frac": valid_frac,
                     "kernels_per_step": launches_per_step},
    }
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(workload, cpu_budget)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="all", choices=["all", "configs", "edges", "blocks", "clutter", "sussman"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "sussman":
        if int(os.environ.get("RANK", "0")) == 0:
            emit(run_sussman(args))
        return

    ctx = Ctx()
    primary = "configs" if args.workload == "all" else args.workload
    steps = args.steps if primary != "clutter" else max(3, min(args.steps, 5 * ctx.world))
    line = run_workload(ctx, args, primary, steps, args.warmup, 10.0)
    if args.workload == "all":
        sec = {}
        for wl in SECONDARY:
            ctx.barrier()
            if wl == "sussman":
                if ctx.rank == 0:
                    try:
                        sec[wl] = run_sussman(args, steps=min(args.steps, 20))
                    except Exception as e:  # the secondary block never takes the headline line down with it
                        sec[wl] = {"metric": metric_name(wl), "error": repr(e)[:300]}
                continue
            if wl == "fk":
                sec[wl] = run_fk(ctx, args, min(args.steps, 50))
                continue
            s_steps = min(args.steps, 5 * ctx.world) if wl == "clutter" else min(args.steps, 50)
            sec[wl] = run_workload(ctx, args, wl, max(3, s_steps), 3, 5.0)
        if ctx.rank == 0:
            line["secondary"] = sec
    ctx.barrier()
    if ctx.rank == 0:
        line["parity"] = PARITY_NOTE
        line["amino_fcl"] = amino_probe()
        emit(line)
    ctx.close()


if __name__ == "__main__":
    main()
