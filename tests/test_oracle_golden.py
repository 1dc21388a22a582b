"""Pins the CPU oracle (oracle/tmk_oracle.c) - no GPU needed.

The reference tree holds no golden vector for the validity check (Makefile.am has
no TESTS), so the oracle is pinned against
  (a) world poses derived by hand from the reference's own scene files
      (demo/baxter-sussman/table.robray, class.robray, sussman-0.robray),
  (b) the static-pair facts of the same scene (Appendix B, computed with an independent SAT),
  (c) the reference's runtime self-checks: demo/tmpexec.c:118 + :424 (after allow_config(q),
      check(q) == 0),
  (d) independent closed forms (point-solid distances, 15-axis SAT) for the GJK primitive,
  (e) the frozen vectors in tests/golden/ (regression).
Parity with Amino/FCL itself stays UNPINNED.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tmkit_b200 import scenes as S
from tmkit_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sussman_golden.npz")


def _q(ax, ang):
    ax = np.asarray(ax, float)
    ax /= np.linalg.norm(ax)
    return np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]])


def _rand_tf(rng, spread=0.5):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return np.concatenate([q, rng.uniform(-spread, spread, size=3)])


def test_appendix_b_world_poses(sussman):
    sc, flat, orc, q0 = sussman
    _, ab = orc.fk(q0)
    want = {
        "table_base": [0, 0, -0.382683432, 0.923879533, 0.1, -0.3, 0.05],
        "front_table": [0, 0, -0.382683432, 0.923879533, 0.524264069, -0.724264069, 0.05],
        "curtain": [0, 0, 0, 1, -0.35, 0, 0],
        "bookshelf": [0, 0, 0, 1, 0, -1.15, 0],
        "lab_front_table": [0, 0, 0, 1, 0.8, -0.2, 0.05],
        "left_table": [0, 0, 0.707106781, 0.707106781, 0.3, 0.8, 0.05],
        "right_table": [0, 0, -0.707106781, 0.707106781, -0.2, -0.7, 0.05],
        "block_a": [0, 0, -0.382683432, 0.923879533, 0.524264069, -0.724264069, 0.1051],
        "block_b": [0, 0, -0.382683432, 0.923879533, 0.347487373, -0.901040764, 0.1051],
        "block_c": [0, 0, -0.382683432, 0.923879533, 0.524264069, -0.724264069, 0.2052],
    }
    for name, tf in want.items():
        np.testing.assert_allclose(ab[orc.frame_id(name)], tf, atol=2e-9, err_msg=name)
    # grasp frame relative to right_endpoint: class.robray:61-65
    i, j = orc.frame_id("right_endpoint"), orc.frame_id("end_effector_grasp")
    assert flat["parent"][j] == i
    np.testing.assert_allclose(flat["E"][j], [0, 1, 0, 0, 0, 0, 0.075])


def test_appendix_b_static_pairs():
    """before any allowance: six outright overlaps, two zero-gap pairs in the BAND, the
    +1e-4 m clearances FREE."""
    sc = S.baxter_sussman()
    orc = O.OracleScene(S.flatten(sc))
    q0 = S.q0_vector(sc)
    _, _, ps = orc.check(orc.fk(q0)[1])
    f = orc.frame_id
    for a, b in [("front_table", "bookshelf"), ("front_table", "lab_front_table"), ("front_table", "right_table"),
                 ("curtain", "bookshelf"), ("curtain", "left_table"), ("curtain", "right_table")]:
        assert ps[f(a), f(b)] == O.HIT, (a, b)
    for a, b in [("lab_front_table", "left_table"), ("lab_front_table", "right_table")]:
        assert ps[f(a), f(b)] == O.BAND, (a, b)
    for a, b in [("front_table", "block_a"), ("front_table", "block_b"), ("block_a", "block_c"),
                 ("lab_front_table", "block_a"), ("right_table", "block_a"), ("right_table", "block_b"),
                 ("block_a", "block_b"), ("block_b", "block_c"), ("left_table", "right_table")]:
        assert ps[f(a), f(b)] == O.FREE, (a, b)
    assert np.array_equal(ps, ps.T)


def test_appendix_b_gaps_with_independent_sat():
    sc = S.baxter_sussman()
    flat = S.flatten(sc)
    orc = O.OracleScene(flat)
    _, ab = orc.fk(S.q0_vector(sc))
    dims = {flat["names"][flat["g_frame"][g]]: flat["g_dim"][g] for g in range(len(flat["g_frame"]))
            if flat["g_shape"][g] == S.BOX}
    gap = lambda a, b: O.sat_box_box(dims[a], ab[orc.frame_id(a)], dims[b], ab[orc.frame_id(b)])
    want = {("front_table", "bookshelf"): -0.0603, ("front_table", "lab_front_table"): -0.0100,
            ("front_table", "right_table"): -0.0100, ("curtain", "bookshelf"): -0.1550,
            ("curtain", "left_table"): -0.1050, ("curtain", "right_table"): -0.6050,
            ("lab_front_table", "left_table"): 0.0, ("lab_front_table", "right_table"): 0.0,
            ("front_table", "block_a"): 1e-4, ("front_table", "block_b"): 1e-4, ("block_a", "block_c"): 1e-4}
    for (a, b), g in want.items():
        assert abs(gap(a, b) - g) < 6e-5, (a, b, gap(a, b))


def test_ka1_allow_config_then_free(sussman):
    """demo/tmpexec.c:118 + :424."""
    sc, flat, orc, q0 = sussman
    st, ex, ps = orc.check(orc.fk(q0)[1])
    assert st == O.FREE and not ex and not ps.any()
    # the 31 listed pairs are in the allowed matrix (allowed-collision.robray:1-31)
    assert len(S.ALLOWED_BAXTER) == 31
    for a, b in S.ALLOWED_BAXTER:
        assert orc.allowed[orc.frame_id(a), orc.frame_id(b)] == 1


def test_q0_file_values():
    """demo/baxter-sussman/q0.tmp:1-2 (mapped by name)."""
    assert S.Q0["right_s0"] == 0.1570796350201586 and S.Q0["left_s0"] == 2.04203514993196
    assert S.Q0["right_w1"] == pytest.approx(np.pi / 2) and S.Q0["left_e1"] == 2.356194490192345
    assert len(S.Q0) == 14


def test_golden_regression(sussman):
    sc, flat, orc, q0 = sussman
    g = np.load(GOLD)
    assert list(g["names"]) == orc.names and list(g["config_names"]) == orc.config_names
    rel, ab = orc.fk(q0)
    np.testing.assert_allclose(ab, g["tf_abs_q0"], atol=1e-14)
    np.testing.assert_allclose(rel, g["tf_rel_q0"], atol=1e-14)
    assert np.array_equal(orc.allowed, g["allowed_after_q0"])
    for q, want in zip(g["fk_q"], g["fk_abs"]):
        np.testing.assert_allclose(orc.fk(q)[1], want, atol=1e-13)
    sub = orc.config_ids(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 4096, 0)
    st, ex = orc.check_batch(sub, q0, Q[:1024], threads=os.cpu_count() or 1)
    assert np.array_equal(st, g["cfg_state_0"][:1024])
    assert np.array_equal(ex, g["cfg_exact_0"][:1024])
    a, b = W.random_edges(lim, 512, 1)
    est, _, first, _ = orc.check_edges(sub, q0, a[:128], b[:128], W.seg_len(lim), threads=os.cpu_count() or 1)
    assert np.array_equal(est, g["edge_state_1"][:128])
    assert np.array_equal(first, g["edge_first_1"][:128])


def test_batch_modes_agree(sussman):
    """hoisting static pairs, early-out and threading do not change verdicts."""
    sc, flat, orc, q0 = sussman
    sub = orc.config_ids(S.RIGHT_ARM)
    Q = W.random_configs(W.arm_limit_table(S.RIGHT_ARM), 300, 5)
    st, ex = orc.check_batch(sub, q0, Q, threads=1)
    st2, ex2 = orc.check_batch(sub, q0, Q, threads=3, hoist=False)
    assert np.array_equal(st, st2) and np.array_equal(ex, ex2)
    st3, ex3 = orc.check_batch(sub, q0, Q, threads=2, early_out=True)
    assert np.array_equal(st3 == O.HIT, st == O.HIT)
    assert np.array_equal(ex3[st != O.BAND], ex[st != O.BAND])
    # and the batch equals the single-configuration restatement of check_state
    for c in range(0, 300, 37):
        q = q0.copy()
        q[sub] = Q[c]
        s1, e1, _ = orc.check(orc.fk(q)[1])
        assert s1 == st[c] and e1 == bool(ex[c])


# ------------------------------------------------------------------ FK restatement
def test_fk_against_matrix_chain():
    """independent 4x4 homogeneous-matrix FK of the right arm."""
    sc = S.baxter_sussman()
    flat = S.flatten(sc)
    orc = O.OracleScene(flat)
    rng = np.random.default_rng(0)

    def mat(tf):
        x, y, z, w = tf[:4]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = tf[4:]
        return M

    for _ in range(5):
        q = rng.uniform(-2, 2, orc.n_q)
        _, ab = orc.fk(q)
        Ms = []
        for i in range(orc.n_f):
            M = mat(flat["E"][i])
            if flat["type"][i] == S.REVOLUTE:
                a = q[flat["qidx"][i]] + flat["offset"][i]
                M = M @ mat(np.concatenate([_q(flat["axis"][i], a), [0, 0, 0]]))
            elif flat["type"][i] == S.PRISMATIC:
                T = np.eye(4)
                T[:3, 3] = flat["axis"][i] * (q[flat["qidx"][i]] + flat["offset"][i])
                M = M @ T
            p = flat["parent"][i]
            Ms.append(M if p < 0 else Ms[p] @ M)
            np.testing.assert_allclose(mat(ab[i]), Ms[i], atol=1e-12)


def test_fk_layout_is_xyzw_then_translation():
    """class.robray:61-65 / lisp/driver.lisp:5: quaternion x,y,z,w then x,y,z."""
    sc = S.Scene()
    sc.add(S.Frame("a", "", S.REVOLUTE, (0, 0, 0, 1), (1, 2, 3), (0, 0, 1.0), "j"))
    orc = O.OracleScene(S.flatten(sc))
    rel, ab = orc.fk(np.array([np.pi / 2]))
    np.testing.assert_allclose(ab[0], [0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4), 1, 2, 3], atol=1e-15)


# ------------------------------------------------------------------ the GJK primitive vs closed forms
@pytest.mark.parametrize("shape", [O.BOX, O.CYLINDER, O.SPHERE])
def test_sphere_vs_solid_closed_form(shape):
    rng = np.random.default_rng(shape)
    n_band = 0
    for k in range(400):
        dim = {O.BOX: rng.uniform(0.05, 0.4, 3), O.CYLINDER: np.array([rng.uniform(0.05, 0.4), rng.uniform(0.02, 0.2), 0]),
               O.SPHERE: np.array([rng.uniform(0.02, 0.3), 0, 0])}[shape]
        tf = _rand_tf(rng, 0.3)
        r = rng.uniform(0.01, 0.2)
        c = rng.uniform(-0.6, 0.6, 3)
        d = O.point_dist(shape, dim, tf, c)
        st, ex = O.pair(O.SPHERE, [r, 0, 0], np.concatenate([[0, 0, 0, 1], c]), shape, dim, tf)
        sep = d - r
        if sep > 2e-6:
            assert st == O.FREE and not ex, (k, sep)
        elif sep < -2e-6 and d > 0:
            assert st == O.HIT and ex, (k, sep)
        else:
            n_band += 1
    assert n_band < 200


def test_box_box_vs_sat():
    rng = np.random.default_rng(7)
    n_hit = n_free = 0
    for k in range(1500):
        da, db = rng.uniform(0.05, 0.5, 3), rng.uniform(0.05, 0.5, 3)
        ta, tb = _rand_tf(rng, 0.35), _rand_tf(rng, 0.35)
        g = O.sat_box_box(da, ta, db, tb)
        st, ex = O.pair(O.BOX, da, ta, O.BOX, db, tb)
        # SAT's largest axis gap is a lower bound on the distance when positive and exact in sign
        if g > 2e-6:
            assert not ex, (k, g)
            n_free += 1
        elif g < -2e-6:
            assert ex and st == O.HIT, (k, g)
            n_hit += 1
    assert n_hit > 100 and n_free > 100


def test_tri_state_band_is_the_contact_tolerance():
    """two unit boxes approaching face to face: FREE above +1e-6, HIT below -1e-6, BAND between."""
    d = [1.0, 1.0, 1.0]
    a = [0, 0, 0, 1, 0, 0, 0]
    for gap, want in [(1e-4, O.FREE), (2e-6, O.FREE), (0.5e-6, O.BAND), (0.0, O.BAND), (-0.5e-6, O.BAND),
                      (-2e-6, O.HIT), (-1e-3, O.HIT)]:
        st, ex = O.pair(O.BOX, d, a, O.BOX, d, [0, 0, 0, 1, 1.0 + gap, 0.1, -0.2])
        assert st == want, (gap, st)
        if gap > 0:
            assert not ex
        if gap < 0:
            assert ex


def test_cylinder_is_base_at_origin_plus_z():
    """Amino cylinders extend in +z from the frame origin (scenes.py header)."""
    cyl = [0.5, 0.1, 0]  # height, radius
    ident = [0, 0, 0, 1, 0, 0, 0]
    ball = lambda z: O.pair(O.SPHERE, [0.05, 0, 0], [0, 0, 0, 1, 0, 0, z], O.CYLINDER, cyl, ident)[1]
    assert ball(0.25) and ball(0.52) and ball(-0.03)
    assert not ball(0.56) and not ball(-0.06)
    assert O.point_dist(O.CYLINDER, cyl, ident, [0, 0, 0.25]) == 0.0
    assert O.point_dist(O.CYLINDER, cyl, ident, [0, 0, -0.1]) == pytest.approx(0.1)


def test_mesh_is_a_surface():
    """FCL BVHModel semantics: a small object wholly inside a closed mesh does not collide."""
    v, t = S._grid_box_mesh((0.5, 0.5, 0.5), (0, 0, 0), 2)
    sc = S.Scene()
    sc.add(S.Frame("shell", "", S.FIXED, geoms=[S.Geom(S.MESH, (0, 0, 0), v, t)]))
    sc.add(S.Frame("probe", "", S.PRISMATIC, (0, 0, 0, 1), (0, 0, 0), (1.0, 0, 0), "x",
                   geoms=[S.Geom(S.SPHERE, (0.1, 0, 0))]))
    orc = O.OracleScene(S.flatten(sc))
    verdict = lambda x: orc.check(orc.fk(np.array([x]))[1])[0]
    assert verdict(0.0) == O.FREE      # inside the shell
    assert verdict(0.45) == O.HIT      # straddling the wall
    assert verdict(0.7) == O.FREE      # outside


# ------------------------------------------------------------------ OMPL discretisation
def test_ompl_order_small_cases():
    assert O.ompl_order(1).tolist() == []
    assert O.ompl_order(2).tolist() == [1]
    assert O.ompl_order(4).tolist() == [2, 1, 3]
    assert O.ompl_order(8).tolist() == [4, 2, 6, 1, 3, 5, 7]
    assert O.ompl_order(7).tolist() == [3, 1, 5, 2, 4, 6]
    for nd in range(2, 200):
        assert sorted(O.ompl_order(nd).tolist()) == list(range(1, nd))


def test_segment_resolution_is_one_percent_of_extent():
    lim = W.arm_limit_table(S.RIGHT_ARM)
    assert W.max_extent(lim) == pytest.approx(12.430, abs=2e-3)
    assert W.seg_len(lim) == pytest.approx(0.1243, abs=1e-4)


def test_edge_semantics(sussman):
    """an edge is invalid iff its end state or an interior state is; first_bad is the position in
    test order; the start state is not tested."""
    sc, flat, orc, q0 = sussman
    sub = orc.config_ids(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    seg = W.seg_len(lim)
    a, b = W.random_edges(lim, 60, 3)
    est, eex, first, nchk = orc.check_edges(sub, q0, a, b, seg)
    for e in range(60):
        nd = max(1, int(np.ceil(np.linalg.norm(b[e] - a[e]) / seg)))
        order = O.ompl_order(nd)
        assert nchk[e] == nd
        states = np.vstack([b[e][None, :]] + [a[e] + (b[e] - a[e]) * (k / nd) for k in order])
        st, ex = orc.check_batch(sub, q0, states)
        if est[e] == O.FREE:
            assert (st == O.FREE).all()
        if est[e] == O.HIT:
            assert (st == O.HIT).any()
            if first[e] >= 0:
                assert first[e] == int(np.argmax(st == O.HIT))
    est2, _, first2, nchk2 = orc.check_edges(sub, q0, a, b, seg, early_out=True)
    assert np.array_equal(est2 == O.HIT, est == O.HIT)
    assert (nchk2 <= nchk).all()


@pytest.mark.parametrize("which", ["sussman", "blocks"])
def test_bounding_sphere_culls_never_change_a_verdict(which):
    """the oracle's fast mode (bounding spheres of geometries, meshes and triangles; shapes built once per
    configuration) against plain brute force: tri-states, exact verdicts and edge results must be identical."""
    sc = S.baxter_sussman() if which == "sussman" else S.blocksworld_scene(8, seed=8)
    flat = S.flatten(sc)
    q0 = S.q0_vector(sc)
    fast, brute = O.OracleScene(flat), O.OracleScene(S.flatten(sc), culls=False)
    for o in (fast, brute):
        o.allow_config(q0)
    assert np.array_equal(fast.allowed, brute.allowed)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    sub = fast.config_ids(S.RIGHT_ARM)
    Q = W.random_configs(lim, 1500, seed=11)
    cores = os.cpu_count() or 1
    for early in (False, True):
        s1, e1 = fast.check_batch(sub, q0, Q, threads=cores, early_out=early)
        s2, e2 = brute.check_batch(sub, q0, Q, threads=cores, early_out=early)
        assert np.array_equal(s1, s2) and np.array_equal(e1, e2)
    a, b = W.random_edges(lim, 150, seed=12)
    r1 = fast.check_edges(sub, q0, a, b, W.seg_len(lim), threads=cores)
    r2 = brute.check_edges(sub, q0, a, b, W.seg_len(lim), threads=cores)
    for x, y in zip(r1, r2):
        assert np.array_equal(x, y)
    # full pair sets (aa_rx_cl_check with a collision set) at a few configurations, allowances dropped
    fresh_fast, fresh_brute = O.OracleScene(S.flatten(sc)), O.OracleScene(S.flatten(sc), culls=False)
    for c in range(4):
        q = q0.copy()
        q[sub] = Q[c]
        p1, p2 = fresh_fast.check(fresh_fast.fk(q)[1]), fresh_brute.check(fresh_brute.fk(q)[1])
        assert p1[0] == p2[0] and p1[1] == p2[1] and np.array_equal(p1[2], p2[2])


def test_clutter_scene_obstacles_do_not_touch_the_robot_at_q0(built):
    """"rejecting any [obstacle] that intersect the robot at q0".  The generator's own
    bounding test (tmkit_b200/scenes.py, device-free) is checked with the oracle's exact pair set at q0, and its FK
    helper against the oracle's FK."""
    from tmkit_b200 import scenes as S
    for seed in (0, 1):
        sc = S.clutter_scene(512, seed=seed)
        orc = O.OracleScene(S.flatten(sc))
        q0 = S.q0_vector(sc)
        _, ab = orc.fk(q0)
        tf = S.world_poses(sc, S.Q0)
        for i, name in enumerate(orc.names):
            r, v = tf[name]
            sgn = 1.0 if np.dot(r, ab[i, :4]) >= 0 else -1.0
            assert np.abs(sgn * r - ab[i, :4]).max() < 1e-12 and np.abs(v - ab[i, 4:]).max() < 1e-12
        _, _, ps = orc.check(ab, want_set=True)
        obs = np.asarray([n.startswith("obstacle_") for n in orc.names])
        assert (ps[np.ix_(obs, ~obs)] != 0).sum() == 0 and (ps[np.ix_(~obs, obs)] != 0).sum() == 0
        assert obs.sum() == 512
    old = S.clutter_scene(512, seed=0, reject_at_q0=False)   # the round-1 scene keeps them
    orc = O.OracleScene(S.flatten(old))
    _, ab = orc.fk(S.q0_vector(old))
    _, _, ps = orc.check(ab, want_set=True)
    obs = np.asarray([n.startswith("obstacle_") for n in orc.names])
    assert (ps[np.ix_(obs, ~obs)] != 0).sum() + (ps[np.ix_(~obs, obs)] != 0).sum() > 0
