"""The device math (tmkit_b200/csrc/dev_math.cuh) compiled for the host (tests/harness) against the
CPU oracle - no GPU needed.  This checks the arithmetic the kernels run, instruction for
instruction apart from FMA contraction:

  * T = double : verdict equals the oracle's exact verdict for every pair outside the band;
  * T = float  : the certificate contract - SEP is never returned for an oracle HIT, HIT never
                 for an oracle FREE, and UNC (-> double re-check) stays rare;
  * ompl_mid   : the closed-form k-th index of OMPL's FIFO bisection equals the queue simulation
                 (SURVEY.md A.5);
  * FK chain   : FP32 error of a 7-joint chain stays far inside the FP32 allowance.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEP, HIT, UNC = 0, 1, 2
K_SPHERE, K_BOX, K_CYL = 0, 1, 2


@pytest.fixture(scope="module")
def H(built):
    L = C.CDLL(os.path.join(ROOT, "tests", "harness", "libdevmath_host.so"))
    for name in ("h_convex_f32", "h_convex_f64"):
        getattr(L, name).argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    for name in ("h_tri_shape_f32", "h_tri_shape_f64"):
        getattr(L, name).argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.h_quick_f32.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    for name in ("h_quick_tri_f32", "h_quick_tri_f64"):
        getattr(L, name).argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.h_tri_tri_f64.argtypes = [C.c_void_p, C.c_void_p]
    L.h_ompl_mid.argtypes = [C.c_int, C.c_int]
    L.h_fk_chain.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return L


def _rand_tf(rng, spread):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return np.concatenate([q, rng.uniform(-spread, spread, size=3)])


def _qrot(q, v):
    u = q[:3]
    t = 2 * np.cross(u, v)
    return v + q[3] * t + np.cross(u, t)


def _rand_shape(rng, kind):
    """returns (device dims, oracle shape id, oracle dims, oracle tf from the centred device tf)"""
    if kind == K_SPHERE:
        r = rng.uniform(0.02, 0.3)
        return np.array([r, 0, 0]), O.SPHERE, np.array([r, 0, 0]), lambda tf: tf
    if kind == K_BOX:
        h = rng.uniform(0.02, 0.4, 3)
        return h, O.BOX, 2 * h, lambda tf: tf
    r, hh = rng.uniform(0.02, 0.2), rng.uniform(0.02, 0.3)

    def to_base(tf):
        out = tf.copy()
        out[4:] = tf[4:] - _qrot(tf[:4], np.array([0, 0, hh]))
        return out
    return np.array([r, hh, 0]), O.CYLINDER, np.array([2 * hh, r, 0]), to_base


PAIRS = [(K_SPHERE, K_SPHERE), (K_SPHERE, K_BOX), (K_SPHERE, K_CYL), (K_BOX, K_BOX), (K_BOX, K_CYL), (K_CYL, K_CYL)]


@pytest.mark.parametrize("ka,kb", PAIRS)
def test_convex_pairs_vs_oracle(H, ka, kb):
    rng = np.random.default_rng(100 * ka + kb)
    n = 3000
    n_unc = n_band = n_hit = n_free = 0
    for i in range(n):
        da, sa, oda, fa = _rand_shape(rng, ka)
        db, sb, odb, fb = _rand_shape(rng, kb)
        ta, tb = _rand_tf(rng, 0.3), _rand_tf(rng, 0.3)
        if i % 3 == 0:  # near-contact cases: slide B along the centre line until the shapes about touch
            d = tb[4:] - ta[4:]
            d /= np.linalg.norm(d) + 1e-30
            lo, hi = 0.0, 3.0
            for _ in range(60):
                mid = 0.5 * (lo + hi)
                t2 = tb.copy()
                t2[4:] = ta[4:] + d * mid
                if O.pair(sa, oda, fa(ta), sb, odb, fb(t2))[1]:
                    lo = mid
                else:
                    hi = mid
            tb[4:] = ta[4:] + d * (hi + rng.choice([-1, 1]) * 10.0 ** rng.uniform(-7, -3))
        st, ex = O.pair(sa, oda, fa(ta), sb, odb, fb(tb))
        r64 = H.h_convex_f64(ka, da.ctypes.data, ta.ctypes.data, kb, db.ctypes.data, tb.ctypes.data)
        r32 = H.h_convex_f32(ka, da.ctypes.data, ta.ctypes.data, kb, db.ctypes.data, tb.ctypes.data)
        assert r64 in (SEP, HIT)
        if st == O.BAND:
            n_band += 1
        else:
            assert (r64 == HIT) == (st == O.HIT), (i, st, r64)
            if r32 == SEP:
                assert st == O.FREE, (i, "float SEP on an oracle HIT")
            elif r32 == HIT:
                assert st == O.HIT, (i, "float HIT on an oracle FREE")
        n_unc += r32 == UNC
        n_hit += st == O.HIT
        n_free += st == O.FREE
    assert n_hit > 200 and n_free > 200
    # a third of the cases are constructed within 1e-7..1e-3 m of contact; away from contact UNC must be rare
    assert n_unc < 0.25 * n, n_unc


NEED = 3


@pytest.mark.parametrize("ka,kb", PAIRS + [(K_CYL, K_BOX), (K_BOX, K_SPHERE), (K_CYL, K_SPHERE)])
def test_quick_certificates_are_sound(H, ka, kb):
    """the warp-uniform quick stage (dev_math.cuh quick_pair): SEP never on an oracle HIT, HIT never on
    an oracle FREE - including flat "table" boxes, thin cylinders and near-contact placements - and
    it decides most pairs by itself."""
    rng = np.random.default_rng(1000 + 10 * ka + kb)
    n = 4000
    n_decided = n_far = 0
    for i in range(n):
        da, sa, oda, fa = _rand_shape(rng, ka)
        db, sb, odb, fb = _rand_shape(rng, kb)
        if i % 4 == 1 and kb == K_BOX:  # a table-like slab
            db = np.array([rng.uniform(0.2, 2.5), rng.uniform(0.2, 2.5), 0.005])
            odb = 2 * db
        if i % 4 == 2 and ka == K_BOX:
            da = np.array([0.005, rng.uniform(0.2, 2.5), rng.uniform(0.2, 2.5)])
            oda = 2 * da
        ta, tb = _rand_tf(rng, 0.5), _rand_tf(rng, 0.5)
        if i % 3 == 0:
            d = tb[4:] - ta[4:]
            d /= np.linalg.norm(d) + 1e-30
            lo, hi = 0.0, 6.0
            for _ in range(60):
                mid = 0.5 * (lo + hi)
                t2 = tb.copy()
                t2[4:] = ta[4:] + d * mid
                if O.pair(sa, oda, fa(ta), sb, odb, fb(t2))[1]:
                    lo = mid
                else:
                    hi = mid
            tb[4:] = ta[4:] + d * (hi + rng.choice([-1, 1]) * 10.0 ** rng.uniform(-7, -2))
        st, ex = O.pair(sa, oda, fa(ta), sb, odb, fb(tb))
        r = H.h_quick_f32(ka, da.ctypes.data, ta.ctypes.data, kb, db.ctypes.data, tb.ctypes.data)
        assert r in (SEP, HIT, UNC, NEED)
        if r == SEP:
            assert st != O.HIT, (i, "quick SEP on an oracle HIT")
        elif r == HIT:
            assert st != O.FREE, (i, "quick HIT on an oracle FREE")
        if i % 3 != 0:
            n_far += 1
            n_decided += r in (SEP, HIT)
    assert n_decided > 0.6 * n_far, (n_decided, n_far)


@pytest.mark.parametrize("kind", [K_SPHERE, K_BOX, K_CYL])
def test_triangle_vs_shape(H, kind):
    rng = np.random.default_rng(kind + 50)
    n_unc = 0
    n = 2000
    for i in range(n):
        d, so, od, f = _rand_shape(rng, kind)
        tri = rng.uniform(-0.5, 0.5, 9)
        ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
        st, ex = O.pair(so, od, f(ident), O.TRI, tri, None)
        r64 = H.h_tri_shape_f64(kind, d.ctypes.data, tri.ctypes.data)
        r32 = H.h_tri_shape_f32(kind, d.ctypes.data, tri.ctypes.data)
        if st != O.BAND:
            assert (r64 == HIT) == (st == O.HIT), (i, st, r64)
            if r32 != UNC:
                assert (r32 == HIT) == (st == O.HIT), (i, st, r32)
        n_unc += r32 == UNC
    assert n_unc < 0.05 * n


@pytest.mark.parametrize("kind", [K_BOX, K_CYL])
def test_quick_triangle_certificates_are_sound(H, kind):
    """quick_tri (mesh leaves): SEP never on an oracle HIT, HIT never on an oracle FREE; in double
    (eps = 0) its verdicts agree with the exact one."""
    rng = np.random.default_rng(kind + 700)
    n = 6000
    n_dec = 0
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
    for i in range(n):
        d, so, od, f = _rand_shape(rng, kind)
        c = rng.uniform(-0.4, 0.4, 3)
        tri = (c[None, :] + rng.uniform(-0.25, 0.25, (3, 3)) * rng.choice([0.05, 0.3, 1.0])).reshape(9)
        if i % 5 == 0:
            tri[[2, 5, 8]] = d[1 if kind == K_CYL else 2] + rng.choice([-1, 1]) * 10.0 ** rng.uniform(-7, -2)  # near a cap / face
        st, ex = O.pair(so, od, f(ident), O.TRI, tri, None)
        r32 = H.h_quick_tri_f32(kind, d.ctypes.data, tri.ctypes.data)
        r64 = H.h_quick_tri_f64(kind, d.ctypes.data, tri.ctypes.data)
        assert r32 in (SEP, HIT, NEED) and r64 in (SEP, HIT, NEED)
        if r32 == SEP:
            assert st != O.HIT, (i, "quick_tri SEP on an oracle HIT")
        if r32 == HIT:
            assert st != O.FREE, (i, "quick_tri HIT on an oracle FREE")
        if st != O.BAND and r64 != NEED:
            assert (r64 == HIT) == (st == O.HIT), (i, st, r64)
        n_dec += r32 != NEED
    assert n_dec > 0.5 * n, n_dec


def test_triangle_triangle_exact(H):
    rng = np.random.default_rng(9)
    n_hit = 0
    for i in range(2000):
        a, b = rng.uniform(-0.4, 0.4, 9), rng.uniform(-0.4, 0.4, 9)
        st, ex = O.pair(O.TRI, a, None, O.TRI, b, None)
        r = H.h_tri_tri_f64(a.ctypes.data, b.ctypes.data)
        if st != O.BAND:
            assert (r == HIT) == (st == O.HIT), i
        n_hit += st == O.HIT
    assert n_hit > 50


def test_ompl_mid_equals_queue_simulation(H):
    for nd in list(range(2, 260)) + [511, 512, 513, 1000, 4097]:
        want = O.ompl_order(nd)
        got = [H.h_ompl_mid(nd, k) for k in range(nd - 1)]
        assert got == want.tolist(), nd


def test_fk_chain_fp32_error(H):
    """7 revolute joints with Baxter-sized offsets: FP32 pose error << the 1e-5 allowance the
    certificates budget for (dev_math.cuh Num<float>::eps)."""
    from tmkit_b200 import scenes as S
    rng = np.random.default_rng(0)
    E = np.array([list(S.quat_from_rpy(*rpy)) + list(xyz) for (_, xyz, rpy, _) in S._ARM_CHAIN])
    axis = np.tile([0.0, 0, 1], (7, 1))
    worst = 0.0
    for _ in range(2000):
        q = rng.uniform(-3, 3, 7).astype(np.float32).astype(np.float64)
        o32, o64 = np.zeros(7), np.zeros(7)
        H.h_fk_chain(7, E.ctypes.data, axis.ctypes.data, q.ctypes.data, 1, o32.ctypes.data)
        H.h_fk_chain(7, E.ctypes.data, axis.ctypes.data, q.ctypes.data, 0, o64.ctypes.data)
        worst = max(worst, np.abs(o32[4:] - o64[4:]).max(), np.abs(o32[:4] - o64[:4]).max())
    assert worst < 2e-6, worst
