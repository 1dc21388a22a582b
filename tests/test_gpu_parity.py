"""GPU parity: the CUDA path, called through the C ABI (ctypes mirror in tmkit_b200.amino),
against the CPU oracle on the same seeded inputs.

Contract (BASELINE.json north_star): collision verdicts bit-exact for every configuration the
oracle classifies FREE (separation > 1e-6 m) or HIT (still colliding after a 1e-6 m erosion);
configurations in the contact BAND are excluded and their count is asserted small.  FK within
1e-5 relative (the FP32 batch FK) / 1e-12 (the double single-configuration ABI).

Reference call shapes being exercised: demo/tmpexec.c:362-428 (check_state), :118 (allow_config),
:307-348 (reparent).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tmkit_b200 import amino as A
from tmkit_b200 import scenes as S
from tmkit_b200 import workloads as W

pytestmark = pytest.mark.gpu

CORES = os.cpu_count() or 1
FK_RTOL = 1e-5  # north_star: FK transforms within 1e-5 relative


def _gpu_scene(sc, q0=None):
    sg = A.SceneGraph.from_scene(sc)
    if q0 is None:
        q0 = np.asarray([S.Q0.get(n, 0.0) for n in sg.config_names()])
    sg.allow_config(q0)
    return sg, A.Collision(sg), q0


def _oracle_scene(sc, q0):
    orc = O.OracleScene(S.flatten(sc))
    orc.allow_config(q0)
    return orc


def _assert_verdicts(valid, st, what, max_band_frac=0.01):
    decided = st != O.BAND
    mism = np.nonzero(((~valid) != (st == O.HIT)) & decided)[0]
    assert mism.size == 0, f"{what}: {mism.size} verdicts differ from the oracle outside the band, first {mism[:8]}"
    assert (~decided).mean() <= max_band_frac, f"{what}: {(~decided).sum()} configurations in the contact band"


def tf_close(a, b, rtol):
    """quaternions compared up to sign; translations relative to the scene scale (1 m)."""
    a, b = np.asarray(a), np.asarray(b)
    qa, qb = a[..., :4], b[..., :4]
    sgn = np.sign(np.sum(qa * qb, axis=-1, keepdims=True))
    sgn[sgn == 0] = 1
    dq = np.abs(qa - sgn * qb).max()
    dv = np.abs(a[..., 4:] - b[..., 4:]).max() / max(1.0, np.abs(b[..., 4:]).max())
    return max(dq, dv) <= rtol, max(dq, dv)


@pytest.fixture(scope="module")
def gpu_sussman(sussman):
    sc, flat, orc, q0 = sussman
    sg, cl, q0g = _gpu_scene(sc, q0)
    return sc, orc, sg, cl, q0


# ------------------------------------------------------------------ FK (aa_rx_sg_tf)
def test_names_and_order(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    assert sg.frame_names() == orc.names
    assert sg.config_names() == orc.config_names
    assert [sg.frame_parent(i) for i in range(sg.frame_count())] == list(orc.flat["parent"])


def test_fk_single_matches_oracle_and_golden(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sussman_golden.npz"))
    rel, ab = sg.tf(q0)
    ok, err = tf_close(ab, g["tf_abs_q0"], 1e-12)
    assert ok, err
    ok, err = tf_close(rel, g["tf_rel_q0"], 1e-12)
    assert ok, err
    for q, want in zip(g["fk_q"], g["fk_abs"]):
        _, ab = sg.tf(q)
        ok, err = tf_close(ab, want, 1e-12)
        assert ok, err


def test_fk_appendix_b_poses(gpu_sussman):
    """world poses derived by hand from the reference's scene files."""
    sc, orc, sg, cl, q0 = gpu_sussman
    _, ab = sg.tf(q0)
    want = {
        "front_table": [0, 0, -0.382683432, 0.923879533, 0.524264069, -0.724264069, 0.05],
        "block_a": [0, 0, -0.382683432, 0.923879533, 0.524264069, -0.724264069, 0.1051],
        "block_b": [0, 0, -0.382683432, 0.923879533, 0.347487373, -0.901040764, 0.1051],
        "block_c": [0, 0, -0.382683432, 0.923879533, 0.524264069, -0.724264069, 0.2052],
        "left_table": [0, 0, 0.707106781, 0.707106781, 0.3, 0.8, 0.05],
        "right_table": [0, 0, -0.707106781, 0.707106781, -0.2, -0.7, 0.05],
    }
    for name, tf in want.items():
        np.testing.assert_allclose(ab[sg.frame_id(name)], tf, atol=2e-9)


def test_fk_batch_fp32(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 300, seed=5)  # ragged: not a multiple of the CTA size
    out = sg.tf_batch(sub, q0, Q)
    assert out.shape == (300, sg.frame_count(), 7)
    worst = 0.0
    for c in range(0, 300, 7):
        q = sg.config_set(sub, Q[c], q0)
        _, ab = orc.fk(q)
        ok, err = tf_close(out[c], ab, FK_RTOL)
        worst = max(worst, err)
        assert ok, (c, err)
    # subset of frames, in caller order
    frames = [sg.frame_id("right_endpoint"), sg.frame_id("block_a"), sg.frame_id("right_s0")]
    out2 = sg.tf_batch(sub, q0, Q, frames=frames)
    np.testing.assert_array_equal(out2, out[:, frames, :])


def test_fk_batch_device_entry_point_all_layouts(gpu_sussman):
    """aa_rx_sg_tf_batch_dev: every input layout x both output layouts agree with the host call and with the oracle
    (FK within 1e-5 relative, north_star); a canary guards the planes' tail"""
    import torch
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    n = 5000 + 13
    Q = W.random_configs(lim, n, seed=6)
    frames = [sg.frame_id(f) for f in ("right_endpoint", "right_w1", "block_a", "right_s0", "right_e1")]
    want = sg.tf_batch(sub, q0, Q, frames=frames)
    for c in range(0, n, 501):
        _, ab = orc.fk(sg.config_set(sub, Q[c], q0))
        ok, err = tf_close(want[c], ab[frames], FK_RTOL)
        assert ok, (c, err)
    ins = {A.Q_ROWS_F64: (torch.from_numpy(Q).cuda(), len(sub)),
           A.Q_ROWS_F32: (torch.from_numpy(Q.astype(np.float32)).cuda(), len(sub)),
           A.Q_SOA_F32: (torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).cuda(), n)}
    s = torch.cuda.current_stream().cuda_stream
    for layout, (dq, ld) in ins.items():
        rows = torch.zeros((n, len(frames), 7), dtype=torch.float64, device="cuda")
        sg.tf_batch_dev(n, sub, q0, dq.data_ptr(), layout, ld, frames, rows.data_ptr(), A.TF_ROWS_F64, 0, s)
        ld_tf = n + 19
        planes = torch.full((len(frames) * 7, ld_tf), 777.0, dtype=torch.float32, device="cuda")
        sg.tf_batch_dev(n, sub, q0, dq.data_ptr(), layout, ld, frames, planes.data_ptr(), A.TF_SOA_F32, ld_tf, s)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(rows.cpu().numpy(), want)
        p = planes.cpu().numpy()
        assert (p[:, n:] == 777.0).all()
        got = p[:, :n].reshape(len(frames), 7, n).transpose(2, 0, 1)
        np.testing.assert_array_equal(got.astype(np.float64), want)


def test_rel_tf_and_reparent(gpu_sussman):
    """pick: re-hang block_c under the grasp frame (demo/tmpexec.c:317-336); world poses must not move."""
    sc, orc, sg, cl, q0 = gpu_sussman
    _, ab = sg.tf(q0)
    i_p, i_c = sg.frame_id("end_effector_grasp"), sg.frame_id("block_c")
    rel = sg.rel_tf(i_p, i_c, ab)
    sg2 = sg.copy()
    sg2.reparent_name("end_effector_grasp", "block_c", rel)
    sg2.init()
    sg2.cl_init()
    assert sg2.frame_count() == sg.frame_count()
    _, ab2 = sg2.tf(q0)
    for name in sg.frame_names():
        ok, err = tf_close(ab2[sg2.frame_id(name)], ab[sg.frame_id(name)], 1e-12)
        assert ok, (name, err)
    # the block now moves with the arm
    q1 = q0.copy()
    q1[sg.config_id("right_s0")] += 0.3
    _, ab3 = sg2.tf(q1)
    assert np.abs(ab3[sg2.frame_id("block_c")][4:] - ab[i_c][4:]).max() > 0.05
    # and the batch path agrees with an oracle built from the same mutated scene
    sc2 = S.baxter_sussman()
    f = sc2.frame("block_c")
    f.parent, f.quat, f.xyz = "end_effector_grasp", tuple(rel[:4]), tuple(rel[4:])
    orc2 = _oracle_scene(sc2, q0)
    sg2.allow_config(q0)
    cl2 = A.Collision(sg2)
    sub2 = sg2.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 2048, seed=11)
    valid = cl2.check_batch(sub2, q0, Q)
    st, _ = orc2.check_batch(orc2.config_ids(S.RIGHT_ARM), q0, Q, threads=CORES)
    _assert_verdicts(valid, st, "reparented scene")


# ------------------------------------------------------------------ aa_rx_cl_check
def test_cl_check_pair_set_at_q0():
    """full-report query before any allowance: the colliding frame pairs equal the oracle's."""
    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    orc = O.OracleScene(S.flatten(sc))
    q0 = S.q0_vector(sc)
    cl = A.Collision(sg)
    _, ab = sg.tf(q0)
    cs = A.CollisionSet(sg)
    cs.clear()
    assert cl.check(ab, cs) != 0
    _, _, ps = orc.check(orc.fk(q0)[1])
    m = cs.matrix()
    assert np.array_equal(m[ps == O.HIT], np.ones((ps == O.HIT).sum(), dtype=np.uint8))
    assert not m[ps == O.FREE].any()
    # the six outright static overlaps of SURVEY.md Appendix B
    for a, b in [("front_table", "bookshelf"), ("front_table", "lab_front_table"), ("front_table", "right_table"),
                 ("curtain", "bookshelf"), ("curtain", "left_table"), ("curtain", "right_table")]:
        assert cs.get(sg.frame_id(a), sg.frame_id(b)) == 1
    # blocks clear what is under them by 1e-4 m
    for a, b in [("front_table", "block_a"), ("front_table", "block_b"), ("block_a", "block_c")]:
        assert cs.get(sg.frame_id(a), sg.frame_id(b)) == 0


def test_ka1_allow_config_then_check_is_free(gpu_sussman):
    """demo/tmpexec.c:118 + :424: after allow_config(q0), check(q0) == 0."""
    sc, orc, sg, cl, q0 = gpu_sussman
    _, ab = sg.tf(q0)
    assert cl.check(ab, None) == 0
    cs = A.CollisionSet(sg)
    assert cl.check(ab, cs) == 0 and cs.pairs() == []
    valid = cl.check_batch(sg.config_indices(S.RIGHT_ARM), q0, q0[sg.config_indices(S.RIGHT_ARM)][None, :])
    assert valid.tolist() == [True]


def test_cl_check_random_configs_match_oracle(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 48, seed=3)
    for c in range(len(Q)):
        q = sg.config_set(sub, Q[c], q0)
        _, ab = sg.tf(q)
        st, ex, ps = orc.check(orc.fk(q)[1])
        cs = A.CollisionSet(sg)
        r = cl.check(ab, cs)
        if st != O.BAND:
            assert (r != 0) == (st == O.HIT), c
        m = cs.matrix()
        assert m[ps == O.HIT].all() and not m[ps == O.FREE].any(), c
        assert (cl.check(ab, None) != 0) == (r != 0)


def test_cl_allow_and_allow_set(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 64, seed=9)
    cl2 = A.Collision(sg)
    for c in range(len(Q)):
        q = sg.config_set(sub, Q[c], q0)
        _, ab = sg.tf(q)
        cs = A.CollisionSet(sg)
        if cl2.check(ab, cs):
            cl2.allow_set(cs)
            assert cl2.check(ab, None) == 0
            i, j = cs.pairs()[0]
            cl2.allow(i, j, 0)
            assert cl2.check(ab, None) != 0
            cl2.allow_name(sg.frame_name(i), sg.frame_name(j), 1)
            assert cl2.check(ab, None) == 0
            return
    pytest.fail("no colliding configuration in the sample")


# ------------------------------------------------------------------ batched validity
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_batch_matches_golden_and_oracle(gpu_sussman, seed):
    sc, orc, sg, cl, q0 = gpu_sussman
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sussman_golden.npz"))
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 4096, seed)
    valid = cl.check_batch(sub, q0, Q)
    _assert_verdicts(valid, g[f"cfg_state_{seed}"], f"golden seed {seed}")
    # float32 rows and the thread-per-configuration kernel give the same bits
    valid32 = cl.check_batch(sub, q0, Q.astype(np.float32))
    assert np.array_equal(valid, valid32)


def test_batch_larger_sample_vs_oracle(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 40000, seed=123)
    valid = cl.check_batch(sub, q0, Q)
    st, _ = orc.check_batch(orc.config_ids(S.RIGHT_ARM), q0, Q, threads=CORES, early_out=True)
    # early-out mode may report HIT where a full scan would also see BAND pairs; verdict unaffected
    _assert_verdicts(valid, st, "40000 configurations")
    assert 0.2 < valid.mean() < 0.8


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 255, 256, 257, 1000])
def test_batch_ragged_sizes(gpu_sussman, n):
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 1000, seed=77)
    full = cl.check_batch(sub, q0, Q)
    part = cl.check_batch(sub, q0, Q[:n].reshape(n, len(sub)))
    assert np.array_equal(part, full[:n])


def test_batch_leading_dimension_and_layouts(gpu_sussman):
    import torch
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    n = 5000
    Q = W.random_configs(lim, n, seed=21)
    want = cl.check_batch(sub, q0, Q)
    wide = np.full((n, 11), np.nan)
    wide[:, :7] = Q
    assert np.array_equal(cl.check_batch(sub, q0, wide), want)
    dev = torch.device("cuda", 0)
    nw = (n + 31) // 32
    for layout, arr, ld in ((A.Q_ROWS_F64, torch.from_numpy(Q), 7), (A.Q_ROWS_F32, torch.from_numpy(Q.astype(np.float32)), 7),
                            (A.Q_SOA_F32, torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))), n)):
        d = arr.to(dev)
        bits = torch.zeros(nw + 1, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        cl.check_batch_dev(n, sub, q0, d.data_ptr(), layout, ld, bits.data_ptr(), s)
        torch.cuda.synchronize()
        got = A.Collision.unpack_bits(bits.cpu().numpy().view(np.uint32), n)
        assert np.array_equal(got, want), layout


def test_kernel_variants_agree(gpu_sussman, monkeypatch):
    """the fused tile kernel and the thread-per-configuration kernel are two schedules of the same
    arithmetic: identical bits."""
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 20000, seed=31)
    want = cl.check_batch(sub, q0, Q)
    monkeypatch.setenv("TMK_NO_TILE", "1")
    cl2 = A.Collision(sg)
    got = cl2.check_batch(sub, q0, Q)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("env", [{"TMK_QCAP": "4"}, {"TMK_QCAP": "4", "TMK_SPILL_CAP": "0"},
                                 {"TMK_QCAP": "4", "TMK_SPILL_CAP": "16", "TMK_UNC_CAP": "64"}, {"TMK_UNC_CAP": "1"},
                                 {"TMK_TASK_CAP": "32"}])
def test_overflow_paths_keep_verdicts(gpu_sussman, monkeypatch, env):
    """full shared-memory queues spill into global memory, a full spill area hands pairs to the exact
    (double) re-check, a full undecided list triggers the brute-force double kernel, and full task stacks of the
    tree-walk kernel (k_mesh_tasks) hand their items to the exact re-check: slower, same bits - configurations
    and edges."""
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 6000, seed=41)
    a, b = W.random_edges(lim, 700, seed=41)
    want = cl.check_batch(sub, q0, Q)
    ewant, efirst = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cl2 = A.Collision(sg)
    assert np.array_equal(cl2.check_batch(sub, q0, Q), want)
    ev, first = cl2.check_edges(sub, q0, a, b, W.seg_len(lim))
    assert np.array_equal(ev, ewant) and np.array_equal(first, efirst)


def test_track_collisions_feedback(gpu_sussman):
    """:track-collisions (lisp/python/tmsmtpy.lisp:59): OR of the colliding frame pairs of a batch."""
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 512, seed=4)
    cl2 = A.Collision(sg)
    cl2.track_collisions(True)
    valid = cl2.check_batch(sub, q0, Q)
    got = cl2.batch_collisions().matrix()
    want_hit = np.zeros_like(got)
    want_any = np.zeros_like(got)
    osub = orc.config_ids(S.RIGHT_ARM)
    for c in range(len(Q)):
        q = q0.copy()
        q[osub] = Q[c]
        st, ex, ps = orc.check(orc.fk(q)[1])
        want_hit |= (ps == O.HIT).astype(np.uint8)
        want_any |= (ps != O.FREE).astype(np.uint8)
    assert got[want_hit == 1].all()
    assert not got[want_any == 0].any()
    assert np.array_equal(valid, cl.check_batch(sub, q0, Q))


def test_stats_counters(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 4096, seed=0)
    cl2 = A.Collision(sg)
    cl2.stats_enable(True)
    valid = cl2.check_batch(sub, q0, Q)
    st = cl2.stats()
    assert st.n_units == 4096 and st.n_states == 4096
    assert st.n_joints == 7 and st.n_moving_geoms == 10
    assert st.n_invalid == int((~valid).sum())
    assert st.n_pairs > 100 and sum(st.n_narrow[:4]) > 0
    assert np.array_equal(valid, cl.check_batch(sub, q0, Q))


# ------------------------------------------------------------------ swept edges
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_edges_match_golden(gpu_sussman, seed):
    sc, orc, sg, cl, q0 = gpu_sussman
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sussman_golden.npz"))
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    a, b = W.random_edges(lim, 512, seed)
    valid, first = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    st, ofirst = g[f"edge_state_{seed}"], g[f"edge_first_{seed}"]
    _assert_verdicts(valid, st, f"edges seed {seed}", max_band_frac=0.02)
    # first invalid state in OMPL's test order; -2 in the oracle = a BAND state precedes the first HIT
    sure = (st != O.BAND) & (ofirst != -2)
    assert np.array_equal(first[sure], ofirst[sure])
    valid2, _ = cl.check_edges(sub, q0, a, b, W.seg_len(lim), want_first=False)
    assert np.array_equal(valid, valid2)


def test_edges_uniform_pairs_and_degenerate(gpu_sussman):
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    osub = orc.config_ids(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    a, b = W.random_edges(lim, 300, seed=5, kind="uniform")  # long edges: several 32-state rounds
    a[:10] = b[:10]  # zero-length edges: nd = 1, only the end state is tested
    valid, first = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    st, _, ofirst, nchk = orc.check_edges(osub, q0, a, b, W.seg_len(lim), threads=CORES)
    _assert_verdicts(valid, st, "uniform edges", max_band_frac=0.03)
    sure = (st != O.BAND) & (ofirst != -2)
    assert np.array_equal(first[sure], ofirst[sure])
    pts = cl.check_batch(sub, q0, b[:10])
    assert np.array_equal(valid[:10], pts)
    assert nchk.max() > 32
    # n = 0
    v0, f0 = cl.check_edges(sub, q0, a[:0], b[:0], W.seg_len(lim))
    assert v0.size == 0


def test_edges_device_soa(gpu_sussman):
    import torch
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    n = 3000
    a, b = W.random_edges(lim, n, seed=8)
    want, wfirst = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    dev = torch.device("cuda", 0)
    da = torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).to(dev)
    db = torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).to(dev)
    bits = torch.zeros((n + 31) // 32 + 1, dtype=torch.int32, device=dev)
    first = torch.zeros(n, dtype=torch.int32, device=dev)
    cl.check_edges_dev(n, sub, q0, da.data_ptr(), db.data_ptr(), A.Q_SOA_F32, n, W.seg_len(lim), bits.data_ptr(),
                       first.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = A.Collision.unpack_bits(bits.cpu().numpy().view(np.uint32), n)
    assert np.array_equal(got, want)
    assert np.array_equal(first.cpu().numpy(), wfirst)


# ------------------------------------------------------------------ other scenes (BASELINE configs[3], [4])
@pytest.mark.parametrize("n_blocks", [8, 16, 32])
def test_blocksworld_scaling_scenes(n_blocks):
    """mesh robot vs box obstacles, one block in the gripper (moving box vs static boxes / meshes)."""
    sc = S.blocksworld_scene(n_blocks, seed=n_blocks)
    q0 = S.q0_vector(sc)
    sg, cl, _ = _gpu_scene(sc, q0)
    orc = _oracle_scene(sc, q0)
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Q = W.random_configs(lim, 3000, seed=n_blocks)
    valid = cl.check_batch(sub, q0, Q)
    st, _ = orc.check_batch(orc.config_ids(S.RIGHT_ARM), q0, Q, threads=CORES, early_out=True)
    _assert_verdicts(valid, st, f"blocksworld {n_blocks}")
    assert 0.02 < valid.mean() < 0.98


@pytest.mark.parametrize("n_prims,reject", [(128, True), (512, True), (512, False)])
def test_dual_arm_clutter_scene(n_prims, reject):
    """14 varied joints, random primitives (boxes / cylinders / spheres); 512 = BASELINE configs[4]
    (20 moving geometries, ~10K candidate pairs: the image is read through L1, not staged).  reject = obstacles
    that touch the robot at q0 are redrawn; without it they stay and the start-state allowance
    handles them (the round-1 scene, kept as a second case)."""
    sc = S.clutter_scene(n_prims, seed=1, reject_at_q0=reject)
    q0 = S.q0_vector(sc)
    sg, cl, _ = _gpu_scene(sc, q0)
    orc = _oracle_scene(sc, q0)
    names = S.RIGHT_ARM + S.LEFT_ARM
    sub = sg.config_indices(names)
    lim = W.arm_limit_table(names)
    Q = W.random_configs(lim, 4000, seed=2)
    valid = cl.check_batch(sub, q0, Q)
    st, _ = orc.check_batch(orc.config_ids(names), q0, Q, threads=CORES, early_out=True)
    _assert_verdicts(valid, st, "dual-arm clutter")
    a, b = W.random_edges(lim, 200, seed=3)
    ev, first = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    est, _, ofirst, _ = orc.check_edges(orc.config_ids(names), q0, a, b, W.seg_len(lim), threads=CORES)
    _assert_verdicts(ev, est, "dual-arm clutter edges", max_band_frac=0.05)


def test_wavefront_pipeline_equals_the_tile_kernel(monkeypatch):
    """large scenes run the wavefront kernels (wave.cuh: FK / broadphase / quick certificates / export as separate
    full-occupancy launches) instead of the tile kernel: same tests, same arithmetic, the same bits - also through
    the host-buffer call (chunked), across its chunk boundary, and when its queues overflow"""
    import torch
    sc = S.clutter_scene(512, seed=0)
    q0 = S.q0_vector(sc)
    names = S.RIGHT_ARM + S.LEFT_ARM
    lim = W.arm_limit_table(names)
    sg, cl, _ = _gpu_scene(sc, q0)
    sub = sg.config_indices(names)
    n = (1 << 20) + 5000 + 17  # more than one wavefront chunk, ragged
    Q = W.random_configs(lim, n, seed=41)
    dq = torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).cuda()
    nw = (n + 31) // 32
    got = {}
    for mode in ("wave", "tile"):
        if mode == "tile":
            monkeypatch.setenv("TMK_NO_WAVE_NOW", "1")
        bits = torch.full((nw + 8,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        for _ in range(3):  # the third call replays the library's graph
            cl.check_batch_dev(n, sub, q0, dq.data_ptr(), A.Q_SOA_F32, n, bits.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        h = bits.cpu().numpy().view(np.uint32)
        assert (h[nw:] == 0x5A5A5A5A).all()
        got[mode] = A.Collision.unpack_bits(h[:nw].copy(), n)
        got[mode + "_host"] = cl.check_batch(sub, q0, Q[:400_000])
    assert np.array_equal(got["wave"], got["tile"])
    assert np.array_equal(got["wave_host"], got["tile_host"]) and np.array_equal(got["wave_host"], got["wave"][:400_000])
    assert 0.01 < got["wave"].mean() < 0.5
    monkeypatch.delenv("TMK_NO_WAVE_NOW")
    orc = _oracle_scene(sc, q0)
    st, _ = orc.check_batch(orc.config_ids(names), q0, Q[:3000], threads=CORES, early_out=True)
    _assert_verdicts(got["wave"][:3000], st, "wavefront pipeline")
    # the broadphase / quick-certificate stages (a condemned state skips the later ones): any stage count, same bits
    for stages in ("1", "2", "5"):
        monkeypatch.setenv("TMK_WAVE_STAGES", stages)
        sg_s, cl_s, _ = _gpu_scene(sc, q0)
        assert np.array_equal(cl_s.check_batch(sub, q0, Q[:300_000]), got["wave"][:300_000]), stages
    monkeypatch.delenv("TMK_WAVE_STAGES")
    # queues too small for the survivors: the overflow flag sends every state through the exact kernel
    monkeypatch.setenv("TMK_WAVE_QCAP", "500")
    assert np.array_equal(cl.check_batch(sub, q0, Q[:6000]), got["wave"][:6000])


def test_static_grid_equals_brute_force(monkeypatch):
    """large scenes put their small static geometry into a uniform grid; the pair-list sweep (grid off)
    must give the same bits - configurations and edges, obstacles of every size class."""
    sc = S.clutter_scene(300, seed=5)
    # a few big obstacles that stay in the brute-force list, one of them spanning many cells
    sc.add(S.Frame("wall", "", S.FIXED, S.quat_from_rpy(0.2, 0.1, 0.7), (0.9, 0.0, 0.4), geoms=[S.Geom(S.BOX, (0.05, 2.0, 1.5))]))
    sc.add(S.Frame("ball", "", S.FIXED, (0, 0, 0, 1), (0.2, 0.9, 0.9), geoms=[S.Geom(S.SPHERE, (0.35, 0, 0))]))
    q0 = S.q0_vector(sc)
    names = S.RIGHT_ARM + S.LEFT_ARM
    lim = W.arm_limit_table(names)
    Q = W.random_configs(lim, 20000, seed=6)
    a, b = W.random_edges(lim, 1500, seed=7)
    sg, cl, _ = _gpu_scene(sc, q0)
    sub = sg.config_indices(names)
    v_grid = cl.check_batch(sub, q0, Q)
    e_grid, f_grid = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    monkeypatch.setenv("TMK_NO_GRID", "1")
    cl2 = A.Collision(sg)
    assert np.array_equal(cl2.check_batch(sub, q0, Q), v_grid)
    e2, f2 = cl2.check_edges(sub, q0, a, b, W.seg_len(lim))
    assert np.array_equal(e2, e_grid) and np.array_equal(f2, f_grid)
    assert 0.005 < v_grid.mean() < 0.9
    orc = _oracle_scene(sc, q0)
    st, _ = orc.check_batch(orc.config_ids(names), q0, Q[:1500], threads=CORES, early_out=True)
    _assert_verdicts(v_grid[:1500], st, "clutter with a grid")


def test_branching_joint_tree():
    """a lift joint carrying two short arms, all varied: the second arm's first joint does not follow its
    parent directly in the joint list, so its parent pose comes from the joint-pose slab."""
    sc = S.Scene()
    sc.add(S.Frame("floor", "", S.FIXED, geoms=[S.Geom(S.BOX, (2.0, 2.0, 0.02))]))
    sc.add(S.Frame("lift", "", S.PRISMATIC, (0, 0, 0, 1), (0, 0, 0.3), (0, 0, 1.0), "lift", 0.0, (0.0, 0.6),
                   geoms=[S.Geom(S.BOX, (0.2, 0.3, 0.2))]))
    for side, y in (("r", -0.2), ("l", 0.2)):
        sc.add(S.Frame(f"{side}0", "lift", S.REVOLUTE, S.quat_from_rpy(0, 0, 0), (0, y, 0.0), (0, 1.0, 0), f"{side}0", 0.0, (-2.5, 2.5)))
        sc.add(S.Frame(f"{side}0c", f"{side}0", S.FIXED, S.quat_from_rpy(0, 1.5708, 0), (0.0, 0, 0), geoms=[S.Geom(S.CYLINDER, (0.3, 0.04, 0))]))
        sc.add(S.Frame(f"{side}1", f"{side}0", S.REVOLUTE, (0, 0, 0, 1), (0.3, 0, 0), (0, 0, 1.0), f"{side}1", 0.0, (-2.5, 2.5)))
        sc.add(S.Frame(f"{side}1c", f"{side}1", S.FIXED, S.quat_from_rpy(0, 1.5708, 0), (0.0, 0, 0), geoms=[S.Geom(S.CYLINDER, (0.25, 0.03, 0))]))
        sc.add(S.Frame(f"{side}tip", f"{side}1", S.FIXED, (0, 0, 0, 1), (0.27, 0, 0), geoms=[S.Geom(S.SPHERE, (0.05, 0, 0))]))
    sc.add(S.Frame("post", "", S.FIXED, (0, 0, 0, 1), (0.35, 0.0, 0.5), geoms=[S.Geom(S.CYLINDER, (0.6, 0.05, 0))]))
    sc.allowed = [("lift", "r0c"), ("lift", "l0c"), ("r0c", "r1c"), ("l0c", "l1c"), ("r1c", "rtip"), ("l1c", "ltip")]
    names = ["lift", "r0", "r1", "l0", "l1"]
    sg = A.SceneGraph.from_scene(sc)
    cl = A.Collision(sg)
    orc = O.OracleScene(S.flatten(sc))
    q0 = np.zeros(5)
    sub = sg.config_indices(names)
    lim = np.asarray([[0.0, 0.6], [-2.5, 2.5], [-2.5, 2.5], [-2.5, 2.5], [-2.5, 2.5]])
    Q = W.random_configs(lim, 8000, seed=3)
    valid = cl.check_batch(sub, q0, Q)
    st, _ = orc.check_batch(orc.config_ids(names), q0, Q, threads=CORES)
    _assert_verdicts(valid, st, "branching tree")
    assert 0.05 < valid.mean() < 0.95
    a, b = W.random_configs(lim, 600, seed=4), W.random_configs(lim, 600, seed=5)
    seg = 0.01 * float(np.linalg.norm(lim[:, 1] - lim[:, 0]))
    ev, first = cl.check_edges(sub, q0, a, b, seg)
    est, _, ofirst, _ = orc.check_edges(orc.config_ids(names), q0, a, b, seg, threads=CORES)
    _assert_verdicts(ev, est, "branching tree edges", max_band_frac=0.05)
    out = sg.tf_batch(sub, q0, Q[:50])
    for c in range(50):
        q = q0.copy()
        q[orc.config_ids(names)] = Q[c]
        ok, err = tf_close(out[c], orc.fk(q)[1], FK_RTOL)
        assert ok, (c, err)


def test_fk_batch_in_a_scene_with_hundreds_of_frames():
    """aa_rx_sg_tf_batch evaluates only the requested frames and their ancestors."""
    sc = S.clutter_scene(512, seed=0)
    sg = A.SceneGraph.from_scene(sc)
    orc = O.OracleScene(S.flatten(sc))
    q0 = S.q0_vector(sc)
    sub = sg.config_indices(S.RIGHT_ARM)
    Q = W.random_configs(W.arm_limit_table(S.RIGHT_ARM), 40, seed=1)
    frames = [sg.frame_id("right_endpoint"), sg.frame_id("obstacle_17"), sg.frame_id("right_e1")]
    out = sg.tf_batch(sub, q0, Q, frames=frames)
    for c in range(40):
        q = sg.config_set(sub, Q[c], q0)
        ab = orc.fk(q)[1]
        ok, err = tf_close(out[c], ab[frames], FK_RTOL)
        assert ok, (c, err)
    full = sg.tf_batch(sub, q0, Q[:4])  # all 570 frames: the slab shrinks its thread count to fit
    assert full.shape == (4, sg.frame_count(), 7)
    ok, err = tf_close(full[0], orc.fk(sg.config_set(sub, Q[0], q0))[1], FK_RTOL)
    assert ok, err


def _random_scene(rng):
    """a random articulated tree (2-6 joints, arbitrary axes and offsets, revolute and prismatic, every
    primitive kind incl. a moving mesh now and then) among random static obstacles incl. a mesh."""
    sc = S.Scene()
    sc.add(S.Frame("base", "", S.FIXED, _rand_quat(rng), tuple(rng.uniform(-0.2, 0.2, 3))))
    n_j = int(rng.integers(2, 7))
    names, lims, links = [], [], ["base"]
    for j in range(n_j):
        parent = links[int(rng.integers(max(0, len(links) - 2), len(links)))]  # mostly a chain, sometimes a branch
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        prismatic = rng.random() < 0.25
        lim = (-0.3, 0.3) if prismatic else (-2.8, 2.8)
        name = f"j{j}"
        sc.add(S.Frame(name, parent, S.PRISMATIC if prismatic else S.REVOLUTE, _rand_quat(rng), tuple(rng.uniform(-0.25, 0.25, 3)),
                       tuple(ax), name, float(rng.uniform(-0.3, 0.3)), lim))
        kind = rng.random()
        if kind < 0.35:
            g = S.Geom(S.BOX, tuple(rng.uniform(0.04, 0.25, 3)))
        elif kind < 0.7:
            g = S.Geom(S.CYLINDER, (float(rng.uniform(0.08, 0.3)), float(rng.uniform(0.02, 0.07)), 0.0))
        elif kind < 0.92:
            g = S.Geom(S.SPHERE, (float(rng.uniform(0.03, 0.09)), 0, 0))
        else:
            v, t = S._grid_box_mesh(tuple(rng.uniform(0.03, 0.1, 3)), (0, 0, 0), 2)
            g = S.Geom(S.MESH, (0, 0, 0), v, t)
        sc.add(S.Frame(f"{name}_g", name, S.FIXED, _rand_quat(rng), tuple(rng.uniform(-0.1, 0.1, 3)), geoms=[g]))
        sc.allowed.append((f"{parent}_g" if parent != "base" else "base", f"{name}_g"))
        names.append(name)
        lims.append(lim)
        links.append(name)
    sc.allowed = [(a, b) for a, b in sc.allowed if a != "base"]
    for k in range(int(rng.integers(4, 14))):
        c = tuple(rng.uniform(-0.8, 0.8, 3))
        kind = rng.random()
        if kind < 0.4:
            g = S.Geom(S.BOX, tuple(rng.uniform(0.03, 0.6, 3)))
        elif kind < 0.7:
            g = S.Geom(S.CYLINDER, (float(rng.uniform(0.05, 0.5)), float(rng.uniform(0.02, 0.12)), 0.0))
        elif kind < 0.9:
            g = S.Geom(S.SPHERE, (float(rng.uniform(0.03, 0.2)), 0, 0))
        else:
            v, t = S._tube_mesh(float(rng.uniform(0.05, 0.2)), float(rng.uniform(0.03, 0.15)), -0.2, 0.2, 8, 2)
            g = S.Geom(S.MESH, (0, 0, 0), v, t)
        sc.add(S.Frame(f"obs{k}", "", S.FIXED, _rand_quat(rng), c, geoms=[g]))
    return sc, names, np.asarray(lims, dtype=np.float64)


def _rand_quat(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return tuple(float(x) for x in q)


@pytest.mark.parametrize("seed", range(24))
def test_random_scenes_fuzz(seed):
    """generic parity: random trees, axes, joint kinds, shapes, meshes - configurations and edges."""
    rng = np.random.default_rng(1000 + seed)
    sc, names, lim = _random_scene(rng)
    sg = A.SceneGraph.from_scene(sc)
    orc = O.OracleScene(S.flatten(sc))
    q0 = np.zeros(len(orc.config_names))
    sg.allow_config(q0)
    orc.allow_config(q0)
    cl = A.Collision(sg)
    sub = sg.config_indices(names)
    osub = orc.config_ids(names)
    Q = W.random_configs(lim, 1500, seed=seed)
    valid = cl.check_batch(sub, q0, Q)
    st, _ = orc.check_batch(osub, q0, Q, threads=CORES)
    _assert_verdicts(valid, st, f"random scene {seed}", max_band_frac=0.03)
    a, b = W.random_configs(lim, 150, seed=seed + 50), W.random_configs(lim, 150, seed=seed + 90)
    seg = 0.01 * float(np.linalg.norm(lim[:, 1] - lim[:, 0]))
    ev, first = cl.check_edges(sub, q0, a, b, seg)
    est, _, ofirst, _ = orc.check_edges(osub, q0, a, b, seg, threads=CORES)
    _assert_verdicts(ev, est, f"random scene {seed} edges", max_band_frac=0.08)
    sure = (est != O.BAND) & (ofirst != -2)
    assert np.array_equal(first[sure], ofirst[sure])
    out = sg.tf_batch(sub, q0, Q[:20])
    for c in range(20):
        q = q0.copy()
        q[osub] = Q[c]
        ok, err = tf_close(out[c], orc.fk(q)[1], FK_RTOL)
        assert ok, (seed, c, err)


def test_prismatic_and_sphere_scene():
    """a hand-built scene with a prismatic joint, spheres and a static mesh."""
    sc = S.Scene()
    sc.add(S.Frame("base", "", S.FIXED, geoms=[S.Geom(S.BOX, (0.2, 0.2, 0.1))]))
    sc.add(S.Frame("slide", "base", S.PRISMATIC, S.quat_from_rpy(0.1, 0.2, 0.3), (0, 0, 0.2), (1.0, 0, 0), "x", 0.05,
                   (-1, 1), geoms=[S.Geom(S.SPHERE, (0.08, 0, 0))]))
    sc.add(S.Frame("turn", "slide", S.REVOLUTE, (0, 0, 0, 1), (0, 0, 0.15), (0, 1.0, 0), "r", 0.0, (-3, 3)))
    sc.add(S.Frame("tip", "turn", S.FIXED, (0, 0, 0, 1), (0.3, 0, 0), geoms=[S.Geom(S.CYLINDER, (0.2, 0.03, 0))]))
    sc.add(S.Frame("ball", "", S.FIXED, (0, 0, 0, 1), (0.5, 0.05, 0.35), geoms=[S.Geom(S.SPHERE, (0.12, 0, 0))]))
    v, t = S._tube_mesh(0.1, 0.05, 0.0, 0.3, 10, 3)
    sc.add(S.Frame("post", "", S.FIXED, (0, 0, 0, 1), (-0.4, 0.0, 0.1), geoms=[S.Geom(S.MESH, (0, 0, 0), v, t)]))
    sc.allowed = [("base", "slide")]
    q0 = np.zeros(2)
    sg = A.SceneGraph.from_scene(sc)
    cl = A.Collision(sg)
    orc = O.OracleScene(S.flatten(sc))
    sub = sg.config_indices(["x", "r"])
    Q = W.random_configs(np.asarray([[-1.0, 1.0], [-3.0, 3.0]]), 5000, seed=1)
    valid = cl.check_batch(sub, q0, Q)
    st, _ = orc.check_batch(orc.config_ids(["x", "r"]), q0, Q, threads=CORES)
    _assert_verdicts(valid, st, "prismatic scene")
    assert 0.05 < valid.mean() < 0.95
    out = sg.tf_batch(sub, q0, Q[:64])
    for c in range(64):
        ok, err = tf_close(out[c], orc.fk(Q[c])[1], FK_RTOL)
        assert ok, (c, err)


# ------------------------------------------------------------------ full-size properties (BASELINE configs[1], [2])
def test_full_size_properties(gpu_sussman):
    """2^20 configurations: determinism, split invariance, permutation invariance, and a sampled
    oracle comparison."""
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    n = 1 << 20
    Q = W.random_configs(lim, n, seed=0)
    v1 = cl.check_batch(sub, q0, Q)
    v2 = cl.check_batch(sub, q0, Q)
    assert np.array_equal(v1, v2)
    k = 333_333
    assert np.array_equal(cl.check_batch(sub, q0, Q[k:k + 100_001]), v1[k:k + 100_001])
    perm = np.random.default_rng(0).permutation(n)
    assert np.array_equal(cl.check_batch(sub, q0, Q[perm]), v1[perm])
    idx = np.random.default_rng(1).choice(n, 20000, replace=False)
    st, _ = orc.check_batch(orc.config_ids(S.RIGHT_ARM), q0, Q[idx], threads=CORES, early_out=True)
    _assert_verdicts(v1[idx], st, "2^20 sample")


def test_full_size_edges_properties(gpu_sussman):
    """100K edges: an edge is valid iff every state of its discretisation is valid (checked through
    the configuration entry point), and reversal keeps interior states."""
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    n = 100_000
    a, b = W.random_edges(lim, n, seed=0)
    seg = W.seg_len(lim)
    valid, first = cl.check_edges(sub, q0, a, b, seg)
    assert np.array_equal(valid, first < 0)
    # end state alone
    vb = cl.check_batch(sub, q0, b)
    assert not valid[~vb].any()
    assert (first[~vb] == 0).all()
    # expand a sample of edges into their states and push them through the batch entry point
    idx = np.random.default_rng(2).choice(n, 2000, replace=False)
    for e in idx[:400]:
        nd = max(1, int(np.ceil(np.linalg.norm(b[e] - a[e]) / seg)))
        order = O.ompl_order(nd)
        ts = np.concatenate([[1.0], order / nd])
        states = a[e][None, :] + (b[e] - a[e])[None, :] * ts[:, None]
        states[0] = b[e]
        sv = cl.check_batch(sub, q0, states)
        if valid[e]:
            assert sv.all(), e
        else:
            assert not sv.all(), e
            assert first[e] == int(np.argmin(sv)), e


def test_context_churn_recycles_buffers_without_stale_state():
    """a task-motion planner creates a collision context per motion query; device buffers, pinned slots and
    streams are recycled through process-wide pools (cl.cu).  Interleave contexts of two scenes of different
    size and check every one against the oracle: a recycled buffer must never leak state into the next user."""
    scenes = []
    for sc in (S.baxter_sussman(), S.blocksworld_scene(16, seed=16)):
        q0 = S.q0_vector(sc)
        scenes.append((sc, q0, _oracle_scene(sc, q0)))
    lim = W.arm_limit_table(S.RIGHT_ARM)
    expected = {}
    for it in range(12):
        k = it % 2
        sc, q0, orc = scenes[k]
        sg, cl, _ = _gpu_scene(sc, q0)
        sub = sg.config_indices(S.RIGHT_ARM)
        n = [700, 33, 4096, 1][it % 4]
        Q = W.random_configs(lim, n, seed=100 + it % 4)
        valid = cl.check_batch(sub, q0, Q)
        key = (k, it % 4)
        if key not in expected:
            st, _ = orc.check_batch(orc.config_ids(S.RIGHT_ARM), q0, Q, threads=CORES, early_out=True)
            _assert_verdicts(valid, st, f"churn {it}", max_band_frac=1.0 if n < 100 else 0.02)
            expected[key] = valid.copy()
        else:
            assert np.array_equal(valid, expected[key]), f"context {it}: verdicts changed on a recycled context"
        a, b = W.random_edges(lim, 64, seed=it)
        v1, f1 = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
        v2, f2 = cl.check_edges(sub, q0, a, b, W.seg_len(lim))  # same arguments: replayed from a CUDA graph
        v3, f3 = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
        assert np.array_equal(v1, v2) and np.array_equal(f1, f2) and np.array_equal(v1, v3) and np.array_equal(f1, f3)
        st = orc.check_edges(orc.config_ids(S.RIGHT_ARM), q0, a, b, W.seg_len(lim), threads=CORES)[0]
        _assert_verdicts(v1, st, f"churn edges {it}", max_band_frac=0.1)
        del cl, sg


def test_two_threads_two_contexts():
    """handles are per-planner, the library is re-entrant across handles (SURVEY.md 8b threading): two host threads,
    each with its own scene graph and collision context, create / use / destroy contexts concurrently (ctypes
    releases the GIL inside the C calls) and must see the verdicts a single thread sees."""
    import threading
    sc = S.baxter_sussman()
    q0 = S.q0_vector(sc)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    Qs = [W.random_configs(lim, 20000, seed=50 + k) for k in range(2)]
    sg0, cl0, _ = _gpu_scene(sc, q0)
    sub = sg0.config_indices(S.RIGHT_ARM)
    expect = [cl0.check_batch(sub, q0, Q) for Q in Qs]
    edges = W.random_edges(lim, 300, seed=3)
    expect_e = cl0.check_edges(sub, q0, edges[0], edges[1], W.seg_len(lim))
    errors = []

    def worker(k):
        try:
            for it in range(6):
                sg, cl, _ = _gpu_scene(S.baxter_sussman(), q0)
                v = cl.check_batch(sub, q0, Qs[k])
                if not np.array_equal(v, expect[k]):
                    errors.append(f"thread {k} iteration {it}: batch verdicts differ")
                ve, fe = cl.check_edges(sub, q0, edges[0], edges[1], W.seg_len(lim))
                if not (np.array_equal(ve, expect_e[0]) and np.array_equal(fe, expect_e[1])):
                    errors.append(f"thread {k} iteration {it}: edge results differ")
                del cl, sg
        except Exception as ex:  # noqa: BLE001 - reported below
            errors.append(f"thread {k}: {ex!r}")

    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_device_pointer_calls_inside_a_callers_cuda_graph(gpu_sussman):
    """the asynchronous entry points only enqueue work on the given stream, so a caller can
    capture them into a CUDA graph of its own (the library then skips its own graph cache) and replay it."""
    import torch
    sc, orc, sg, cl, q0 = gpu_sussman
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    n = 40000
    Q = W.random_configs(lim, n, seed=77)
    expect = cl.check_batch(sub, q0, Q)
    a, b = W.random_edges(lim, 2000, seed=78)
    exp_v, exp_f = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    dq = torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).cuda()
    da = torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).cuda()
    db = torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).cuda()
    bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    ebits = torch.zeros((2000 + 31) // 32, dtype=torch.int32, device="cuda")
    efirst = torch.zeros(2000, dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()

    def calls():
        cl.check_batch_dev(n, sub, q0, dq.data_ptr(), A.Q_SOA_F32, n, bits.data_ptr(), s.cuda_stream)
        cl.check_edges_dev(2000, sub, q0, da.data_ptr(), db.data_ptr(), A.Q_SOA_F32, 2000, W.seg_len(lim),
                           ebits.data_ptr(), efirst.data_ptr(), s.cuda_stream)

    with torch.cuda.stream(s):
        calls()  # warm-up: scratch buffers of these sizes exist afterwards
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s, capture_error_mode="relaxed"):
            calls()
        for _ in range(3):
            bits.zero_(); ebits.zero_(); efirst.zero_()
            g.replay()
            s.synchronize()
            got = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
            assert np.array_equal(got, expect)
            gv = np.unpackbits(ebits.cpu().numpy().view(np.uint8), bitorder="little")[:2000].astype(bool)
            assert np.array_equal(gv, exp_v)
            assert np.array_equal(efirst.cpu().numpy(), exp_f)
