#!/usr/bin/env python
"""bench.py - headline benchmark: collision-checked configurations / s (BASELINE.json configs[1]).

A step = one pass of the validity check (FK + broadphase + narrowphase + exact re-check of the
contact band) over one batch of 2^20 random Baxter right-arm configurations against the Sussman
block scene.  `value` is measured with the inputs resident in HBM; `e2e` goes through the
host-buffer C-ABI call (aa_rx_cl_check_batch) with the H2D / D2H copies inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload configs|edges]

N > 1 is launched by torchrun (one rank per GPU); every rank checks its own batch (weak scaling)
and the validity bitmaps are all-gathered over NCCL inside the timed region (second stream, overlapping
the next step; the timed stream waits for the last gathers before the closing event).

The timed steps call the asynchronous device-pointer entry points, which replay a CUDA graph for arguments
they have seen before (the untimed warm-up goes twice around the ring of input batches); the per-kernel
times in `roofline` come from a second, event-instrumented loop.  `cpu_baseline` / `--impl reference` time
the oracle's fast mode (oracle/) on the host cores - the only use of oracle/ here.
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
RING = 8  # distinct input batches cycled through (RING * batch bytes > L2)


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


# workload -> (scene builder, varied joints, units per step per GPU, kind, BASELINE.json config, description)
def workload_table():
    from tmkit_b200 import scenes as S
    return {
        "configs": (lambda: S.baxter_sussman(), S.RIGHT_ARM, N_CFG, "configs",
                    "BASELINE configs[1]: 2^20 random Baxter 7-DOF right-arm configurations vs the Sussman block scene "
                    "(restated Baxter fixture), per GPU"),
        "edges": (lambda: S.baxter_sussman(), S.RIGHT_ARM, N_EDGE, "edges",
                  "BASELINE configs[2]: 100K random RRT-like Baxter right-arm segments, OMPL 1% resolution, per GPU"),
        "blocks": (lambda: S.blocksworld_scene(32, seed=32), S.RIGHT_ARM, N_CFG, "configs",
                   "BASELINE configs[3]: 2^20 right-arm configurations, 32-block tabletop, one block in the gripper, "
                   "mesh robot vs box obstacles, per GPU"),
        "clutter": (lambda: S.clutter_scene(512, seed=0), S.RIGHT_ARM + S.LEFT_ARM, 1 << 21, "configs",
                    "BASELINE configs[4]: 2^21 dual-arm (14-DOF) configurations vs 512 random primitives, per GPU "
                    "(16M over 8 GPUs)"),
    }


def make_problem(workload="configs"):
    from tmkit_b200 import amino as A
    from tmkit_b200 import scenes as S
    from tmkit_b200 import workloads as W
    build, names, n_units, kind, desc = workload_table()[workload]
    return build(), S, W, A, names, n_units, kind, desc


def cpu_baseline(workload: str, seconds_budget: float = 12.0):
    """The CPU restatement (oracle, early-out mode = aa_rx_cl_check with a NULL set) on all host
    cores, on a bounded sample of the same workload."""
    from oracle import oracle as O
    sc, S, W, _, names, n_units_w, kind, _ = make_problem(workload)
    so = "/tmp/libtmkoracle_native.so"
    try:
        path = O.build(force=True, march_native=True, out=so)
    except Exception:
        path = O.build()
    flat = S.flatten(sc)
    orc = O.OracleScene(flat, libpath=path)
    q0 = np.asarray([S.Q0.get(n, 0.0) for n in orc.config_names])
    orc.allow_config(q0)
    sub = orc.config_ids(names)
    lim = W.arm_limit_table(names)
    cores = os.cpu_count() or 1
    if kind == "edges":
        n = 2000
        a, b = W.random_edges(lim, n, seed=0)
        t0 = time.perf_counter()
        orc.check_edges(sub, q0, a, b, W.seg_len(lim), early_out=True, threads=cores)
        dt = time.perf_counter() - t0
        n2 = int(min(200_000, max(n, n * seconds_budget / max(dt, 1e-6))))
        a, b = W.random_edges(lim, n2, seed=1)
        t0 = time.perf_counter()
        orc.check_edges(sub, q0, a, b, W.seg_len(lim), early_out=True, threads=cores)
        dt = time.perf_counter() - t0
        return n2 / dt, cores, f"{n2} of {N_EDGE} RRT-like edges, early-out, all cores", n2, dt
    n = 4000
    Q = W.random_configs(lim, n, seed=0)
    t0 = time.perf_counter()
    orc.check_batch(sub, q0, Q, early_out=True, threads=cores)
    dt = time.perf_counter() - t0
    n2 = int(min(4_000_000, max(n, n * seconds_budget / max(dt, 1e-6))))
    Q = W.random_configs(lim, n2, seed=1)
    t0 = time.perf_counter()
    orc.check_batch(sub, q0, Q, early_out=True, threads=cores)
    dt = time.perf_counter() - t0
    return n2 / dt, cores, f"{n2} of {n_units_w} configurations, early-out, all cores", n2, dt


def run_reference(args):
    """The reference arm: the CPU restatement on all host cores, exactly K timed steps after W warm-up steps, each
    step a bounded sample of the workload (the whole run is sized to about two minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    sc, S, W, _, names, n_units_w, kind, _ = make_problem(args.workload)
    try:
        path = O.build(force=True, march_native=True, out="/tmp/libtmkoracle_native.so")
    except Exception:
        path = O.build()
    orc = O.OracleScene(S.flatten(sc), libpath=path)
    q0 = np.asarray([S.Q0.get(n, 0.0) for n in orc.config_names])
    orc.allow_config(q0)
    sub = orc.config_ids(names)
    lim = W.arm_limit_table(names)
    cores = os.cpu_count() or 1
    edges = kind == "edges"
    unit = "edges/s" if edges else "configs/s"

    def make(n, seed):
        return W.random_edges(lim, n, seed=seed) if edges else (W.random_configs(lim, n, seed=seed),)

    def run(batch):
        t0 = time.perf_counter()
        if edges:
            orc.check_edges(sub, q0, batch[0], batch[1], W.seg_len(lim), early_out=True, threads=cores)
        else:
            orc.check_batch(sub, q0, batch[0], early_out=True, threads=cores)
        return time.perf_counter() - t0

    n_cal = 2000 if edges else 4000
    rate = n_cal / max(run(make(n_cal, 0)), 1e-6)                      # calibration
    budget = min(8.0, max(0.25, 120.0 / max(1, args.steps)))           # seconds per step
    if os.environ.get("TMK_BENCH_REF_BUDGET"):                        # test aid
        budget = float(os.environ["TMK_BENCH_REF_BUDGET"])
    n_step = int(max(n_cal, min(200_000 if edges else 4_000_000, rate * budget)))
    ring = [make(n_step, 1 + r) for r in range(2)]
    for w in range(args.warmup):
        run(ring[w % 2])
    times = [run(ring[i % 2]) for i in range(args.steps)]
    total = float(sum(times))
    v = n_step * args.steps / total
    sample = (f"{n_step} of {n_units_w} {'RRT-like edges' if edges else 'configurations'} per step, early-out, "
              f"bounding-sphere culls, all cores")
    line = {
        "impl": "reference", "sample_units": n_step, "sample_s": total / args.steps, "metric": metric_name(args.workload),
        "value": v, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload),
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of Amino/FCL semantics (the reference's Amino/FCL/OMPL are not in the tree)",
    }
    emit(line)


def metric_name(workload):
    return {"configs": "collision-checked configs/sec (FK + collision, Baxter right arm vs Sussman scene)",
            "edges": "swept-edge checks/sec (RRT-Connect segments at OMPL resolution)",
            "blocks": "collision-checked configs/sec (FK + collision, Baxter right arm, 32-block tabletop)",
            "clutter": "collision-checked configs/sec (FK + collision, dual-arm Baxter vs 512 primitives)"}[workload]


def config_dict(workload):
    _, _, n_units, kind, desc = workload_table()[workload]
    return {"workload": desc, "units_per_step_per_gpu": n_units,
            "l2": f"inputs rotate over a ring of {RING} distinct batches" + (" (> L2)" if kind == "configs" else "")}


def run_sussman(args):
    """BASELINE configs[0]: the motion side of the Baxter Sussman task-motion plan, end to end
    (tmkit_b200/sussman.py), beside the CPU restatement replaying exactly the same validity work."""
    import torch
    from oracle import oracle as O
    from tmkit_b200 import scenes as S
    from tmkit_b200 import sussman as SU
    from tmkit_b200 import workloads as W
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the validity check has no CPU fallback")
    for w in range(max(1, args.warmup)):
        SU.run(seed=900 + w)
    runs = []
    for k in range(args.steps):
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
            orc = O.OracleScene(S.flatten(desc))
            orc.allow_config(q_start)
            lim = W.arm_limit_table(names)
            t0 = time.perf_counter()
            orc.check_edges(orc.config_ids(names), q_start, q0e, q1e, W.seg_len(lim), early_out=True, threads=th)
            t += time.perf_counter() - t0
            if label == "1":
                n_edges += len(q0e)
        cpu[label] = t
    split = {k: float(np.median([x.times[k] for x in ok])) for k in ("ik", "rrt", "simplify", "scene")} if ok else {}
    line = {
        "metric": "Baxter Sussman motion-plan time (6 motion queries: IK + RRT-Connect + simplify + reparent)",
        "value": tot, "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
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
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="configs", choices=["configs", "edges", "blocks", "clutter", "sussman"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.workload == "sussman":
        run_sussman(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the validity check has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    sc, S, W, A, names, n_units, kind, _ = make_problem(args.workload)
    sg = A.SceneGraph.from_scene(sc)
    q0 = np.asarray([S.Q0.get(n, 0.0) for n in sg.config_names()])
    sg.allow_config(q0)  # the reference allows what collides at the start configuration
    cl = A.Collision(sg)
    sub = sg.config_indices(names)
    lim = W.arm_limit_table(names)
    seg = W.seg_len(lim)
    edges = kind == "edges"
    unit = "edges/s" if edges else "configs/s"
    nw = (n_units + 31) // 32

    # ---- inputs: a ring of distinct batches, SoA float32 in HBM (device-resident leg) ----
    ring, ring1 = [], []
    host_rows, host_rows1 = [], []
    for r in range(RING):
        seed = 1000 * rank + r
        if edges:
            a, b = W.random_edges(lim, n_units, seed=seed)
            ring.append(torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).to(dev))
            ring1.append(torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).to(dev))
            if r < 2:
                host_rows.append(torch.from_numpy(a).pin_memory()); host_rows1.append(torch.from_numpy(b).pin_memory())
        else:
            Q = W.random_configs(lim, n_units, seed=seed)
            ring.append(torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).to(dev))
            if r < 2:
                host_rows.append(torch.from_numpy(Q).pin_memory())
    # two result buffers: the all-gather of step k (communication stream) overlaps the kernels of step k+1
    bits2 = [torch.zeros(nw, dtype=torch.int32, device=dev) for _ in range(2)]
    bits = bits2[0]
    gathered2 = [torch.zeros(nw * world, dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None
    # a dedicated stream: the C ABI takes a cudaStream_t (NULL would mean the handle's own stream),
    # and the events below must be recorded on the stream the kernels are launched on
    ts = torch.cuda.Stream(device=dev)
    stream = ts.cuda_stream
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # high priority: when the persistent tile kernel of the next step and the all-gather of this one are both
    # pending, the collective gets the first free SMs (it is tiny and everybody waits for it)
    cs = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    ev_done = [torch.cuda.Event() for _ in range(2)]      # kernels of the step wrote bits2[b]
    ev_gathered = [torch.cuda.Event() for _ in range(2)]  # the all-gather has read bits2[b]

    sub_c = np.ascontiguousarray(sub, dtype=np.int32)
    q0_c = np.ascontiguousarray(q0, dtype=np.float64)
    sub_ptr, q0_ptr, n_sub_c, n_q_c, seg_c = sub_c.ctypes.data, q0_c.ctypes.data, len(sub_c), len(q0_c), float(seg)
    ring_ptr = [t.data_ptr() for t in ring]
    ring1_ptr = [t.data_ptr() for t in ring1]
    bits_ptr = [t.data_ptr() for t in bits2]
    c_batch, c_edges = cl.L.aa_rx_cl_check_batch_dev, cl.L.aa_rx_cl_check_edges_dev

    def step(i):
        k = i % RING
        b = i & 1 if world > 1 else 0
        out = bits2[b]
        if world > 1:
            ts.wait_event(ev_gathered[b])  # the gather of step i-2 is done with this buffer
        # the C entry points directly, arguments marshalled once: the loop below has to keep a ~1 ms GPU step fed
        if edges:
            c_edges(cl.h, n_units, n_sub_c, sub_ptr, n_q_c, q0_ptr, ring_ptr[k], ring1_ptr[k], A.Q_SOA_F32, n_units, seg_c,
                    bits_ptr[b], None, stream)
        else:
            c_batch(cl.h, n_units, n_sub_c, sub_ptr, n_q_c, q0_ptr, ring_ptr[k], A.Q_SOA_F32, n_units, bits_ptr[b], stream)
        if world > 1:
            ev_done[b].record(ts)
            cs.wait_event(ev_done[b])
            with torch.cuda.stream(cs):
                dist.all_gather_into_tensor(gathered2[b], out)
                ev_gathered[b].record(cs)

    def drain():
        """the timed stream waits for the gathers still in flight: they belong to the timed steps"""
        if world > 1:
            ts.wait_event(ev_gathered[0])
            ts.wait_event(ev_gathered[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(ts):
        l2_flush.zero_()
        # the device-pointer entry points capture a CUDA graph the second time they see the same arguments: the
        # untimed warm-up goes twice around the ring so that the timed steps replay graphs (steady state)
        n_warm = max(args.warmup, 2 * RING)
        for i in range(n_warm):
            step(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        # events around the dominant kernel, recorded by the library on `stream`.  The timed steps are
        # replayed from CUDA graphs (no events inside a graph): the kernel times come from a second, untimed loop
        inline_timing = False
        cl.kernel_timing(inline_timing)
        barrier()
        ev[0].record(ts)
        t_host = time.perf_counter()
        for i in range(args.steps):
            step(n_warm + i)
        drain()
        ev[1].record(ts)
        host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps
        if rank == 0:
            # graph replays put the host far ahead of the GPU: the clocks are sampled now, with everything
            # enqueued and the GPU still working through the timed steps (an NVML query inside the loop would
            # stall the enqueueing thread for milliseconds)
            sampler.sample_now()
        barrier()
        ms_total = ev[0].elapsed_time(ev[1])
        if not inline_timing:
            cl.kernel_timing(True)
            for i in range(args.steps):
                step(n_warm + i)
            barrier()
        n_timed, k_ms_sum, rk_ms_sum = cl.kernel_time()
        cl.kernel_timing(False)
        clocks = sampler.stop() if rank == 0 else None
    assert n_timed == args.steps, (n_timed, args.steps)
    k_ms = k_ms_sum / n_timed
    rk_ms = rk_ms_sum / n_timed

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # per-rank step time, kernel time and host enqueue time (diagnostics beside the max the contract asks for)
        mine = torch.tensor([ms_total / args.steps, k_ms + rk_ms, host_enqueue_ms], dtype=torch.float64, device=dev)
        allr = torch.zeros(3 * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = [[round(float(x), 4) for x in allr[3 * r:3 * r + 3].tolist()] for r in range(world)]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_units * world / (ms_step * 1e-3)
    valid_frac = float(np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[:n_units].mean())

    # ---- e2e: host buffers through the public C-ABI call (H2D + kernels + D2H inside) ----
    e2e_steps = max(3, min(args.steps, 10))
    hb = np.zeros(nw + 1, dtype=np.uint32)

    def e2e_step(i):
        k = i % len(host_rows)
        if edges:
            cl.L.aa_rx_cl_check_edges(cl.h, n_units, len(sub), sub.ctypes.data, len(q0), q0.ctypes.data,
                                      host_rows[k].data_ptr(), host_rows1[k].data_ptr(), len(sub), seg,
                                      hb.ctypes.data, None)
        else:
            cl.L.aa_rx_cl_check_batch(cl.h, n_units, len(sub), sub.ctypes.data, len(q0), q0.ctypes.data,
                                      host_rows[k].data_ptr(), len(sub), hb.ctypes.data)

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)  # blocking: returns with the bitmap in host memory
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_units * world * e2e_steps / float(t.item())
    h2d = n_units * len(sub) * 8 * (2 if edges else 1)
    d2h = nw * 4
    # beside it (information only): the same blocking call with float rows - half the PCIe bytes, which is what
    # bounds the double-row call (the reference-facing one quoted as `e2e`)
    e2e_f32 = None
    if not edges and world == 1:
        rows32 = [torch.from_numpy(np.ascontiguousarray(h.numpy().astype(np.float32))).pin_memory() for h in host_rows]
        def f32_step(i):
            cl.L.aa_rx_cl_check_batch_f32(cl.h, n_units, len(sub), sub.ctypes.data, len(q0), q0.ctypes.data,
                                          rows32[i % len(rows32)].data_ptr(), len(sub), hb.ctypes.data)
        for i in range(3):
            f32_step(i)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            f32_step(i)
        e2e_f32 = n_units * e2e_steps / (time.perf_counter() - t0)

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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_kind, sm_max = load_peaks()
    traffic = None  # dram bytes per launch of the dominant kernel, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(args.workload)
        if tj:
            traffic = tj["bytes_per_launch"] * (n_units / tj["units_per_launch"])
    except Exception:
        traffic = None
    bytes_unit = (4 * len(sub) * (2 if edges else 1)) + 1.0 / 8
    achieved_gbs = bytes_unit * n_units / (k_ms * 1e-3) / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    fp32_probe = cl.L.tmk_fp32_peak_tflops()  # measured on this device: dependent-free FFMA loop, all SMs
    fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    flops_unit = f_cfg * states_per_unit
    achieved_tf = flops_unit * n_units / (k_ms * 1e-3) / 1e12
    kernel = ("k_states_tile<SrcEdgeLevel> + k_heavy_mixed + k_edges_compact, all bisection levels" if edges
              else "k_states_tile<SrcConfigs> + deferred narrowphase (k_conv_items, k_mesh_traverse, k_tri_items)")

    line = {
        "metric": metric_name(args.workload), "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(config_dict(args.workload), parallelism=f"dp{world}: contiguous slice per rank" + (
            ", one NCCL all-gather of the validity words per step on a second stream (overlaps the next step; "
            "all gathers finish inside the timed region)" if world > 1 else "")),
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "aa_rx_cl_check_edges" if edges else "aa_rx_cl_check_batch", "host_dtype": "f64 rows, pinned",
                "value_f32_rows": e2e_f32},
        "gpu_launches": int(launches_per_step * args.steps),
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
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic, "algorithmic_bytes": bytes_unit * n_units,
                         "peak_kind": peak_kind, "kernel": kernel,
                         "kernel_ms": k_ms, "bytes_per_unit": bytes_unit},
        "counters": {"joints": int(st.n_joints), "moving_geoms": int(st.n_moving_geoms), "pairs": int(st.n_pairs),
                     "narrow_per_state": [st.n_narrow[k] / max(1, st.n_states) for k in range(4)],
                     "states_per_unit_one_warp_per_edge_kernel": states_per_unit_ref,
                     "double_rechecked_pairs_per_state": unc_frac, "valid_frac": valid_frac,
                     "kernels_per_step": launches_per_step},
    }
    if not args.no_cpu:
        v, cores, sample, n, dt = cpu_baseline(args.workload)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                                "note": "CPU restatement of Amino/FCL semantics (oracle, early-out mode)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
