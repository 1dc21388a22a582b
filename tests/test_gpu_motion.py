"""The caller of the path: workspace-goal IK, RRT-Connect on the batched edge
entry point, shortcut simplification, reparenting, plan-file output - on the Baxter Sussman problem
(reference demo/baxter-sussman; refinement semantics demo/domains/blocksworld/tm-blocks.py:200-266).
Every returned path is re-validated against the CPU oracle at OMPL resolution."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tmkit_b200 import amino as A
from tmkit_b200 import scenes as S
from tmkit_b200 import sussman as SU
from tmkit_b200 import workloads as W

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 1


def _validate_path(desc, path, names, goal=None):
    orc = O.OracleScene(S.flatten(desc))
    assert orc.config_names == desc.config_names()
    orc.allow_config(path[0])  # start-state allowance
    sub = orc.config_ids(names)
    lim = W.arm_limit_table(names)
    assert (path[:, sub] >= lim[:, 0] - 1e-9).all() and (path[:, sub] <= lim[:, 1] + 1e-9).all()
    others = np.setdiff1d(np.arange(path.shape[1]), sub)
    assert np.array_equal(path[:, others], np.repeat(path[:1, others], len(path), axis=0))
    st, ex, first, _ = orc.check_edges(sub, path[0], path[:-1, sub], path[1:, sub], W.seg_len(lim), threads=CORES)
    assert (st != O.HIT).all(), f"path edge {np.nonzero(st == O.HIT)[0]} collides according to the oracle"
    return orc


def test_joint_goal_query(built):
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q0 = S.q0_vector(sc)
    ssg = A.SubSceneGraph(sg, "", SU.FRAME)
    assert ssg.config_names() == S.RIGHT_ARM
    mp = A.MotionPlanner(ssg)
    mp.set_simplify(True)
    mp.set_start(q0)
    goal = q0[ssg.config_ids()].copy()
    goal[0] -= 1.2  # swing the shoulder: blocked by nothing or by the table, either way a path exists
    goal[1] -= 0.3
    assert mp.set_goal(goal) == 1
    path = mp.plan(3.0)
    assert path is not None and len(path) >= 2
    np.testing.assert_array_equal(path[0], q0)
    np.testing.assert_allclose(path[-1][ssg.config_ids()], goal)
    _validate_path(sc, path, S.RIGHT_ARM)
    # an invalid goal is refused: arm through the table
    bad = q0[ssg.config_ids()].copy()
    bad[1] = 1.0
    bad[3] = 0.0
    if mp.set_goal(bad) == 0:
        mp2 = A.MotionPlanner(ssg)
        mp2.set_start(q0)
        assert mp2.set_goal(bad) == 0 and mp2.plan(0.2) is None  # no goal -> planning failure


def test_workspace_goal_ik(built):
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q0 = S.q0_vector(sc)
    _, ab = sg.tf(q0)
    ssg = A.SubSceneGraph(sg, "", SU.FRAME)
    mp = A.MotionPlanner(ssg)
    mp.set_start(q0)
    goal = ab[sg.frame_id("block_c")]
    assert mp.set_wsgoal(goal) >= 1
    path = mp.plan(3.0)
    assert path is not None
    _, ab1 = sg.tf(path[-1])
    tip = ab1[sg.frame_id(SU.FRAME)]
    assert np.abs(tip[4:] - goal[4:]).max() < 1e-6
    assert min(np.abs(tip[:4] - goal[:4]).max(), np.abs(tip[:4] + goal[:4]).max()) < 1e-6
    # unreachable goal: 3 m away
    far = goal.copy()
    far[4] += 3.0
    mp.set_start(q0)
    assert mp.set_wsgoal(far) == 0 and mp.plan(0.2) is None


@pytest.mark.parametrize("seed", [0, 1])
def test_sussman_end_to_end(built, tmp_path, seed):
    r = SU.run(seed=seed)
    assert r.ok, f"planning failure at {r.failed_action}"
    assert len(r.paths) == 6
    for k, (path, desc, (tip, goal)) in enumerate(zip(r.paths, r.scene_descs, r.goals)):
        orc = _validate_path(desc, path, S.RIGHT_ARM)
        if k:
            np.testing.assert_array_equal(path[0], r.paths[k - 1][-1])  # each query starts where the last one ended
        tipf = orc.fk(path[-1])[1][orc.frame_id(tip)]
        assert np.abs(tipf[4:] - goal[4:]).max() < 1e-6, (k, tip)
    # goal of the anomaly: A on B on C, each 1e-4 m above what carries it, all above the table
    orc = O.OracleScene(S.flatten(r.scene_final))
    ab = orc.fk(r.q_final)[1]
    za, zb, zc = (ab[orc.frame_id(n)][6] for n in ("block_a", "block_b", "block_c"))
    assert zb - zc == pytest.approx(0.1001, abs=1e-6) and za - zb == pytest.approx(0.1001, abs=1e-6)
    for n in ("block_a", "block_b"):
        assert np.abs(ab[orc.frame_id(n)][4:6] - ab[orc.frame_id("block_c")][4:6]).max() < 1e-6
    assert orc.flat["parent"][orc.frame_id("block_a")] == orc.frame_id("block_b")
    assert orc.flat["parent"][orc.frame_id("block_b")] == orc.frame_id("block_c")
    assert orc.flat["parent"][orc.frame_id("block_c")] == orc.frame_id("front_table")
    # plan file: 6 x (action, motion, reparent), replayable
    f = tmp_path / "sussman.tmp"
    assert r.plan.write(str(f)) == 0
    ops = A.PlanFile.parse(str(f)).ops()
    assert [o[0] for o in ops] == ["a", "m", "r"] * 6
    assert ops[0][1] == "UNSTACK block_c block_a" and ops[2] == ("r", "block_c", SU.FRAME)
    assert ops[1][1] == S.RIGHT_ARM
    for k in range(6):
        np.testing.assert_array_equal(ops[3 * k + 1][2], r.paths[k][:, orc.config_ids(S.RIGHT_ARM)])
    assert r.times["total"] < 6 * SU.MOTION_TIMEOUT


def test_collision_feedback(built):
    """:track-collisions - the frame pairs that blocked extensions (lisp/itmp-rec.lisp:230-247)."""
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q0 = S.q0_vector(sc)
    ssg = A.SubSceneGraph(sg, "", SU.FRAME)
    ids = ssg.config_ids()
    lim = W.arm_limit_table(S.RIGHT_ARM)
    sg.allow_config(q0)
    cl = A.Collision(sg)
    cand = W.random_configs(lim, 400, seed=12)
    cand = cand[cl.check_batch(ids, q0, cand)]
    for goal in cand:
        mp = A.MotionPlanner(ssg)
        mp.set_track_collisions(True)
        mp.set_start(q0)
        assert mp.set_goal(goal) == 1
        path = mp.plan(0.5)
        if mp.stats().n_iterations == 0:
            continue  # straight line was free: nothing was blocked
        # (a sampled goal may lie in another component of the free space, e.g. under the table:
        # then the query times out, which is the planning-failure the task planner reacts to)
        pairs = mp.collisions().pairs()
        assert len(pairs) > 0
        names = {sg.frame_name(i) for p in pairs for i in p}
        assert any(n.startswith("right_") for n in names)
        return
    pytest.fail("every sampled goal was directly reachable")


def test_collision_feedback_reports_the_first_invalid_state_of_an_edge(built):
    """ADVICE r1 / VERDICT N4: an edge whose END point is free but which is blocked at an interior state must still
    contribute the frame pairs colliding there.  With a zero time budget the planner only checks start -> goal,
    so the feedback is exactly the oracle's pair set at that edge's first invalid state."""
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q0 = S.q0_vector(sc)
    ssg = A.SubSceneGraph(sg, "", SU.FRAME)
    ids = ssg.config_ids()
    lim = W.arm_limit_table(S.RIGHT_ARM)
    seg = W.seg_len(lim)
    sg2 = A.SceneGraph.from_scene(sc)
    sg2.allow_config(q0)
    cl = A.Collision(sg2)
    cand = W.random_configs(lim, 600, seed=14)
    cand = cand[cl.check_batch(ids, q0, cand)]
    start = np.repeat(q0[ids][None, :], len(cand), axis=0)
    valid, first = cl.check_edges(ids, q0, start, cand, seg)
    blocked = np.nonzero(~valid & (first > 0))[0]
    assert blocked.size >= 3, "no sampled goal is free with a blocked straight line"
    orc = O.OracleScene(S.flatten(sc))
    orc.allow_config(q0)
    osub = orc.config_ids(S.RIGHT_ARM)
    checked = 0
    for e in blocked[:6]:
        goal = cand[e]
        est, _, efirst, _ = orc.check_edges(osub, q0, start[e:e + 1], goal[None, :], seg, threads=1)
        if est[0] != O.HIT or efirst[0] != first[e]:
            continue  # a contact-band state on the edge: not a clean case
        mp = A.MotionPlanner(ssg)
        mp.set_track_collisions(True)
        mp.set_start(q0)
        assert mp.set_goal(goal) == 1
        assert mp.plan(0.0) is None and mp.stats().n_iterations == 0
        got = set(mp.collisions().pairs())
        assert got, "an edge blocked at an interior state contributed nothing"
        # the oracle's pair set at that state
        nd = max(1, int(np.ceil(np.linalg.norm(goal - start[e]) / seg)))
        t = O.ompl_order(nd)[first[e] - 1] / nd
        q = q0.copy()
        q[ids] = start[e] + (goal - start[e]) * t
        _, ab = orc.fk(q)
        st, _, ps = orc.check(ab, want_set=True)
        want = {(i, j) for i in range(len(ps)) for j in range(i) if ps[i, j] == O.HIT or ps[j, i] == O.HIT}
        band = {(i, j) for i in range(len(ps)) for j in range(i) if ps[i, j] == O.BAND or ps[j, i] == O.BAND}
        assert want and want <= got <= (want | band), (want, got)
        checked += 1
    assert checked >= 1


def test_simplifier_never_lengthens_and_stays_valid(built):
    """:simplify t = OMPL's reduceVertices -> shortcutPath order (mp.c simplify_path): for the same query and seed the
    simplified path is valid, ends where the raw one ends, and is no longer than it"""
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q0 = S.q0_vector(sc)
    ssg = A.SubSceneGraph(sg, "", SU.FRAME)
    ids = ssg.config_ids()
    lim = W.arm_limit_table(S.RIGHT_ARM)
    sg2 = A.SceneGraph.from_scene(sc)
    sg2.allow_config(q0)
    cl = A.Collision(sg2)
    cand = W.random_configs(lim, 300, seed=15)
    cand = cand[cl.check_batch(ids, q0, cand)]
    start = np.repeat(q0[ids][None, :], len(cand), axis=0)
    valid, _ = cl.check_edges(ids, q0, start, cand, W.seg_len(lim))
    done = 0
    for goal in cand[~valid][:12]:
        paths = []
        for simplify in (False, True):
            mp = A.MotionPlanner(ssg)
            mp.set_seed(5)
            mp.set_simplify(simplify)
            mp.set_start(q0)
            assert mp.set_goal(goal) == 1
            paths.append(mp.plan(1.0))
        raw, simp = paths
        if raw is None or simp is None:
            continue  # the goal lies in another component of the free space
        length = lambda p: float(np.linalg.norm(np.diff(p[:, ids], axis=0), axis=1).sum())
        np.testing.assert_array_equal(simp[0], raw[0])
        np.testing.assert_array_equal(simp[-1], raw[-1])
        assert length(simp) <= length(raw) + 1e-9
        assert length(simp) >= np.linalg.norm(goal - q0[ids]) - 1e-9
        _validate_path(sc, simp, S.RIGHT_ARM)
        done += 1
    assert done >= 3


def test_c_executor_replays_the_plan(built, tmp_path):
    """the whole boundary in C: a scene plugin generated as C source (what the reference's aarxc emits,
    Makefile.am:213-231), loaded with aa_rx_dl_sg, the start file and the plan file read with
    tmk_plan_parse, every waypoint checked with aa_rx_sg_tf + aa_rx_cl_check exactly as
    demo/tmpexec.c:362-428 does, the segments through aa_rx_cl_check_edges at the planner's resolution, reparents applied
    with aa_rx_sg_rel_tf / _copy / _reparent_name / _init (tmpexec.c:307-348)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "harness", "plan_check")
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "tests", "harness"), "plan_check"])
    sc = S.baxter_sussman()
    src = tmp_path / "sussman_scene.c"
    src.write_text(S.to_plugin_source(sc, "sussman"))
    so = tmp_path / "libsussman_scene.so"
    libdir = os.path.join(root, "tmkit_b200")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", f"-I{root}/include", "-o", str(so), str(src), f"-L{libdir}",
                           "-l:libtmkcl.so", f"-Wl,-rpath,{libdir}"])
    r = SU.run(seed=3)
    assert r.ok
    plan_file = tmp_path / "plan.tmp"
    assert r.plan.write(str(plan_file)) == 0
    start = A.PlanFile()
    q0 = S.q0_vector(sc)
    start.add_motion(sc.config_names(), q0[None, :])
    start_file = tmp_path / "q0.tmp"
    assert start.write(str(start_file)) == 0
    p = subprocess.run([exe, str(so), "sussman", str(start_file), str(plan_file), "0.01"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.strip().splitlines()
    assert sum(l.startswith("motion") for l in lines) == 6 and sum(l.startswith("reparent") for l in lines) == 6
    assert lines[-1].startswith("summary  6 motions, 6 reparents") and lines[-1].endswith(" 0 colliding")
    final = {l.split()[1]: np.array(l.split()[2:], dtype=float) for l in lines if l.startswith("final")}
    # the C executor, starting from the plugin's scene, ends with the same stack the Python driver built
    orc = O.OracleScene(S.flatten(r.scene_final))
    ab = orc.fk(r.q_final)[1]
    for name, pos in final.items():
        np.testing.assert_allclose(pos, ab[orc.frame_id(name)][4:], atol=1e-8)
    assert final["block_a"][2] - final["block_b"][2] == pytest.approx(0.1001, abs=1e-6)


def test_workspace_goal_ik_prismatic_chain(built):
    """the one-launch IK (ik.cuh) on a chain that mixes prismatic and revolute joints with tilted, non-unit axes:
    goals generated from random configurations must be reached to 1e-6."""
    sc = S.Scene()
    sc.add(S.Frame("base", "", S.FIXED, geoms=[S.Geom(S.BOX, (0.2, 0.2, 0.1))]))
    sc.add(S.Frame("slide", "base", S.PRISMATIC, S.quat_from_rpy(0.1, 0.2, 0.3), (0, 0, 0.2), (1.0, 0, 0), "x", 0.05, (-1, 1)))
    sc.add(S.Frame("yaw", "slide", S.REVOLUTE, (0, 0, 0, 1), (0, 0, 0.15), (0, 0, 1.0), "a", 0.0, (-3, 3)))
    sc.add(S.Frame("pitch", "yaw", S.REVOLUTE, S.quat_from_rpy(0.3, 0.0, 0.1), (0.2, 0, 0), (0, 2.0, 0), "b", 0.1, (-2, 2)))
    sc.add(S.Frame("lift", "pitch", S.PRISMATIC, (0, 0, 0, 1), (0.1, 0, 0), (0, 0.6, 0.8), "z", 0.0, (-0.5, 0.5)))
    sc.add(S.Frame("roll", "lift", S.REVOLUTE, (0, 0, 0, 1), (0.1, 0, 0), (1.0, 0, 0), "c", 0.0, (-3, 3)))
    sc.add(S.Frame("wrist", "roll", S.REVOLUTE, (0, 0, 0, 1), (0.1, 0, 0), (0, 1.0, 0), "d", 0.0, (-3, 3)))
    sc.add(S.Frame("tool", "wrist", S.FIXED, S.quat_from_rpy(0.0, 0.4, 0.0), (0.15, 0, 0.02),
                   geoms=[S.Geom(S.SPHERE, (0.02, 0, 0))]))
    sc.allowed = [("base", "tool")]
    sg = A.SceneGraph.from_scene(sc)
    names = sg.config_names()
    q0 = np.zeros(len(names))
    rng = np.random.default_rng(5)
    lo = np.array([-1, -3, -2, -0.5, -3, -3.0])
    hi = -lo
    order = {n: i for i, n in enumerate(names)}
    n_found = 0
    for trial in range(8):
        q_star = np.zeros(len(names))
        for n, l, h in zip(["x", "a", "b", "z", "c", "d"], lo, hi):
            q_star[order[n]] = rng.uniform(0.8 * l, 0.8 * h)
        _, ab = sg.tf(q_star)
        goal = ab[sg.frame_id("tool")]
        ssg = A.SubSceneGraph(sg, "", "tool")
        mp = A.MotionPlanner(ssg)
        mp.set_seed(trial)
        mp.set_start(q0)
        if mp.set_wsgoal(goal) == 0:
            continue  # every converged candidate was in collision or none converged: allowed, but rare
        path = mp.plan(3.0)
        if path is None:
            continue
        n_found += 1
        _, ab1 = sg.tf(path[-1])
        tip = ab1[sg.frame_id("tool")]
        assert np.abs(tip[4:] - goal[4:]).max() < 1e-6
        assert min(np.abs(tip[:4] - goal[:4]).max(), np.abs(tip[:4] + goal[:4]).max()) < 1e-6
    assert n_found >= 6, n_found
