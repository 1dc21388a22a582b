"""Baxter Sussman-anomaly demo (BASELINE.json configs[0], reference demo/baxter-sussman): the motion
side of the task-motion plan, end to end.

The task planner (PDDL -> SMT, lisp/planner.lisp) is out of scope; the six-action
task plan it finds for this problem is written down here.  Every action is refined exactly as the
reference's domain script does (demo/domains/blocksworld/tm-blocks.py:200-266):

    PICK-UP / UNSTACK obj     motion of `end_effector_grasp` to obj's pose, then reparent obj to it
    PUT-DOWN obj dst i j      motion of obj to dst * (i*RES, j*RES, h_obj/2 + h_dst/2 + EPSILON),
    STACK obj dst             motion of obj to dst * (0, 0, h_obj/2 + h_dst/2 + EPSILON), then
                              reparent obj to dst

with the motion query of lisp/python/tmsmtpy.lisp:55-60 (workspace goal, :simplify t, timeout 3 s)
answered by aa_rx_mp_* (tmkit_b200/csrc/mp.c) on top of the batched CUDA validity check, and the
scene mutated between queries with aa_rx_sg_copy / aa_rx_sg_reparent_name / aa_rx_sg_init
(demo/tmpexec.c:317-336).  The result is a plan file in the reference's format.
"""
from __future__ import annotations

import copy
import time
from typing import List, Optional, Tuple

import numpy as np

from . import amino as A
from . import scenes as S

FRAME = "end_effector_grasp"  # class.robray:61-65
RESOLUTION = 0.1  # tm-blocks.py:6
EPSILON = 1e-4  # tm-blocks.py:12
MOTION_TIMEOUT = 3.0  # lisp/motion-plan.lisp:3

# the task plan of the Sussman anomaly (C on A, B on the table -> A on B on C).  The put-down cell is
# the one symbolic choice in it; when its motion query fails the next cell of PUT_DOWN_CELLS is tried,
# a one-line stand-in for the constraint-and-replan loop of lisp/itmp-rec.lisp:224-265.
PUT_DOWN_CELLS = [(-2, 0), (-2, 1), (-1, 2), (-2, -1), (1, 2)]
TASK_PLAN = [
    ("UNSTACK", "block_c", "block_a"),
    ("PUT-DOWN", "block_c", "front_table", -2, 0),
    ("PICK-UP", "block_b", "front_table"),
    ("STACK", "block_b", "block_c"),
    ("PICK-UP", "block_a", "front_table"),
    ("STACK", "block_a", "block_b"),
]


def tf_mul(a, b):
    x1, y1, z1, w1 = a[:4]
    x2, y2, z2, w2 = b[:4]
    r = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                  w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    u = np.asarray(a[:3])
    t = 2 * np.cross(u, b[4:])
    return np.concatenate([r, np.asarray(b[4:]) + a[3] * t + np.cross(u, t) + np.asarray(a[4:])])


def box_height(sc: S.Scene, name: str) -> float:
    """place_height (tm-blocks.py:213-217): half the z extent of the frame's box."""
    g = sc.frame(name).geoms[0]
    assert g.shape == S.BOX
    return g.dims[2] / 2


class Result:
    def __init__(self):
        self.ok = False
        self.failed_action: Optional[tuple] = None
        self.plan = A.PlanFile()
        self.paths: List[np.ndarray] = []
        self.scenes: List[A.SceneGraph] = []  # scene graph each motion was planned in
        self.scene_descs: List[S.Scene] = []  # ... and its neutral description (for the oracle)
        self.goals: List[Tuple[str, np.ndarray]] = []  # (tip frame, goal pose) of each motion
        self.q_final: Optional[np.ndarray] = None
        self.sg_final: Optional[A.SceneGraph] = None
        self.times = {"total": 0.0, "ik": 0.0, "rrt": 0.0, "simplify": 0.0, "scene": 0.0, "bookkeeping": 0.0}
        self.counters = {"edges_checked": 0, "edge_calls": 0, "iterations": 0, "ik_iterations": 0, "nodes": 0}
        self.queries: List[dict] = []


def run(seed: int = 0, task_plan=TASK_PLAN, timeout: float = MOTION_TIMEOUT, simplify: bool = True,
        edge_log: Optional[list] = None) -> Result:
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q = S.q0_vector(sc)  # -q q0.tmp (lisp/driver.lisp:92-95)
    res = Result()
    t0 = time.perf_counter()
    sg.tf(q)  # first device call: CUDA context creation, excluded from the plan time like process start-up
    res.times["cuda_init"] = time.perf_counter() - t0
    t_total = time.perf_counter()
    todo = [(act, 0) for act in task_plan]
    k = -1
    while todo:
        act, attempt = todo.pop(0)
        k += 1
        name, obj, dst = act[0], act[1], act[2]
        t0 = time.perf_counter()
        _, ab = sg.tf(q)
        if name in ("PICK-UP", "UNSTACK"):
            tip, goal = FRAME, ab[sg.frame_id(obj)]
        else:
            x, y = (act[3] * RESOLUTION, act[4] * RESOLUTION) if name == "PUT-DOWN" else (0.0, 0.0)
            z = box_height(sc, obj) + box_height(sc, dst) + EPSILON
            tip, goal = obj, tf_mul(ab[sg.frame_id(dst)], np.array([0, 0, 0, 1.0, x, y, z]))
        ssg = A.SubSceneGraph(sg, "", tip)
        mp = A.MotionPlanner(ssg)
        mp.set_seed(1000 * seed + k)
        mp.set_simplify(simplify)
        mp.set_start(q)
        if edge_log is not None:
            mp.edge_log(True)
        res.times["scene"] += time.perf_counter() - t0
        n_goal = mp.set_wsgoal(goal)
        path = mp.plan(timeout) if n_goal else None
        st = mp.stats()
        res.times["ik"] += st.ik_s; res.times["rrt"] += st.rrt_s; res.times["simplify"] += st.simplify_s
        res.counters["edges_checked"] += st.n_edges_checked; res.counters["edge_calls"] += st.n_edge_calls
        res.counters["iterations"] += st.n_iterations; res.counters["ik_iterations"] += st.n_ik_iterations
        res.counters["nodes"] += st.n_nodes
        res.queries.append({"action": " ".join(str(a) for a in act), "goals": int(st.n_goals), "ik_s": st.ik_s, "rrt_s": st.rrt_s,
                            "simplify_s": st.simplify_s, "iterations": int(st.n_iterations),
                            "edges": int(st.n_edges_checked), "waypoints": 0 if path is None else len(path)})
        if edge_log is not None:
            t_book = time.perf_counter()
            edge_log.append((copy.deepcopy(sc), q.copy(), ssg.config_names(), *mp.edge_log_get()))
            res.times["bookkeeping"] += time.perf_counter() - t_book
        if path is None and name == "PUT-DOWN" and attempt + 1 < len(PUT_DOWN_CELLS):
            i, j = PUT_DOWN_CELLS[attempt + 1]
            todo.insert(0, ((name, obj, dst, i, j), attempt + 1))
            continue
        if path is None:
            res.failed_action = act  # -> planning-failure (lisp/tm-plan.lisp:3-7)
            res.times["total"] = time.perf_counter() - t_total - res.times["bookkeeping"]
            return res
        names = ssg.config_names()
        ids = ssg.config_ids()
        res.plan.add_action(" ".join(str(a) for a in act))
        res.plan.add_motion(names, path[:, ids])
        res.paths.append(path)
        res.scenes.append(sg)
        t_book = time.perf_counter()
        res.scene_descs.append(copy.deepcopy(sc))  # bookkeeping for the tests, not part of the plan time
        res.goals.append((tip, np.asarray(goal, dtype=np.float64)))
        res.times["bookkeeping"] += time.perf_counter() - t_book
        q = path[-1].copy()
        # reparent (lisp/tm-plan.lisp:50-59, demo/tmpexec.c:317-336)
        t0 = time.perf_counter()
        _, ab = sg.tf(q)
        new_parent = FRAME if name in ("PICK-UP", "UNSTACK") else dst
        rel = sg.rel_tf(sg.frame_id(new_parent), sg.frame_id(obj), ab)
        sg2 = sg.copy()
        sg2.reparent_name(new_parent, obj, rel)
        sg2.init()
        sg2.cl_init()
        res.plan.add_reparent(obj, new_parent)
        f = sc.frame(obj)
        f.parent, f.quat, f.xyz = new_parent, tuple(rel[:4]), tuple(rel[4:])
        sg = sg2
        res.times["scene"] += time.perf_counter() - t0
    res.ok = True
    res.q_final, res.sg_final = q, sg
    res.times["total"] = time.perf_counter() - t_total - res.times["bookkeeping"]
    res.scene_final = sc
    return res


if __name__ == "__main__":
    import json
    import sys
    r = run(seed=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    print(json.dumps({"ok": r.ok, "times": r.times, "counters": r.counters, "queries": r.queries}, indent=1))
    if r.ok and len(sys.argv) > 2:
        r.plan.write(sys.argv[2])
