"""ctypes wrapper around oracle/libtmkoracle.so.

TEST INFRASTRUCTURE ONLY (see tmk_oracle.c).  Importable from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FREE, HIT, BAND = 0, 1, 2
SPHERE, BOX, CYLINDER, MESH, TRI = 0, 1, 2, 3, 4
TOL = 1e-6  # contact tolerance of the parity contract (BASELINE.json north_star)


class _Scene(C.Structure):
    _fields_ = [
        ("n_f", C.c_int), ("n_q", C.c_int),
        ("parent", C.c_void_p), ("type", C.c_void_p), ("E", C.c_void_p), ("axis", C.c_void_p),
        ("offset", C.c_void_p), ("qidx", C.c_void_p),
        ("n_g", C.c_int),
        ("g_frame", C.c_void_p), ("g_shape", C.c_void_p), ("g_dim", C.c_void_p), ("g_mesh", C.c_void_p),
        ("n_mesh", C.c_int),
        ("mesh_vstart", C.c_void_p), ("mesh_tstart", C.c_void_p), ("verts", C.c_void_p), ("tris", C.c_void_p),
        ("allowed", C.c_void_p),
        ("mesh_bc", C.c_void_p), ("mesh_br", C.c_void_p), ("tri_c", C.c_void_p), ("tri_b", C.c_void_p),
    ]


def build(force: bool = False, march_native: bool = False, out: Optional[str] = None) -> str:
    """Compile the oracle with gcc.  march_native builds a host-tuned copy (used by the
    CPU-baseline timing on the box it runs on)."""
    src = os.path.join(HERE, "tmk_oracle.c")
    so = out or os.path.join(HERE, "libtmkoracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        flags = ["-O3", "-fPIC", "-std=c99", "-fno-fast-math", "-ffp-contract=off"]
        if march_native:
            flags.append("-march=native")
        subprocess.check_call(["gcc", *flags, "-shared", "-o", so, src, "-lm", "-lpthread"])
    return so


_lib = None


def lib(path: Optional[str] = None):
    global _lib
    if _lib is None or path is not None:
        L = C.CDLL(path or build())
        L.orc_fk.argtypes = [C.c_void_p] * 4
        L.orc_check.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_check.restype = C.c_int
        L.orc_check_batch.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                      C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_check_edges.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_long, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pair.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_pair.restype = C.c_int
        L.orc_sat_box_box.argtypes = [C.c_void_p] * 4
        L.orc_sat_box_box.restype = C.c_double
        L.orc_point_dist.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_point_dist.restype = C.c_double
        L.orc_ompl_order.argtypes = [C.c_int, C.c_void_p]
        L.orc_ompl_order.restype = C.c_int
        L.orc_segment_count.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_segment_count.restype = C.c_int
        L.orc_gjk_iters.restype = C.c_long
        if path is not None:
            return L
        _lib = L
    return _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


class OracleScene:
    """Flat scene arrays (from tmkit_b200.scenes.flatten) held alive for the C side."""

    def __init__(self, flat: Dict[str, np.ndarray], libpath: Optional[str] = None, culls: bool = True):
        """culls=False leaves the optional bounding-sphere tables out (plain brute force): only used by the test
        that checks the culls never change a verdict."""
        self.culls = culls
        self.L = lib(libpath) if libpath else lib()
        self.flat = flat
        self.names = list(flat["names"])
        self.config_names = list(flat["config_names"])
        self.n_f = len(self.names)
        self.n_q = len(self.config_names)
        k = self._keep = {}
        for key, dt in (("parent", np.int32), ("type", np.int32), ("E", np.float64), ("axis", np.float64),
                        ("offset", np.float64), ("qidx", np.int32), ("g_frame", np.int32), ("g_shape", np.int32),
                        ("g_dim", np.float64), ("g_mesh", np.int32), ("mesh_vstart", np.int32),
                        ("mesh_tstart", np.int32), ("verts", np.float64), ("tris", np.int32), ("allowed", np.uint8)):
            k[key] = np.ascontiguousarray(flat[key], dtype=dt)
        self.n_g = len(k["g_frame"])
        # bounding sphere of every mesh in its own frame (centre of the vertex box, farthest vertex): lets the C
        # side skip triangles whose outcome is FREE anyway - verdicts are unaffected
        vs, n_mesh = k["mesh_vstart"], len(k["mesh_vstart"]) - 1
        bc = np.zeros((max(n_mesh, 1), 3))
        br = np.zeros(max(n_mesh, 1))
        for m in range(n_mesh):
            v = k["verts"].reshape(-1, 3)[vs[m]:vs[m + 1]]
            bc[m] = 0.5 * (v.min(axis=0) + v.max(axis=0))
            br[m] = np.sqrt(((v - bc[m]) ** 2).sum(axis=1).max()) * (1 + 1e-12)
        k["mesh_bc"], k["mesh_br"] = np.ascontiguousarray(bc), np.ascontiguousarray(br)
        # ... and centroid / radius of every triangle in its mesh frame
        ts = k["mesh_tstart"]
        tri = k["tris"].reshape(-1, 3)
        tc = np.zeros((max(len(tri), 1), 3))
        tb = np.zeros(max(len(tri), 1))
        for m in range(n_mesh):
            v = k["verts"].reshape(-1, 3)[vs[m]:vs[m + 1]]
            p = v[tri[ts[m]:ts[m + 1]]]  # nt x 3 x 3
            c = p.mean(axis=1)
            tc[ts[m]:ts[m + 1]] = c
            tb[ts[m]:ts[m + 1]] = np.sqrt(((p - c[:, None, :]) ** 2).sum(axis=2).max(axis=1)) * (1 + 1e-12)
        k["tri_c"], k["tri_b"] = np.ascontiguousarray(tc), np.ascontiguousarray(tb)
        self._refresh()

    def _refresh(self):
        k = self._keep
        s = _Scene()
        s.n_f, s.n_q, s.n_g = self.n_f, self.n_q, self.n_g
        s.n_mesh = len(k["mesh_vstart"]) - 1
        for key in ("parent", "type", "E", "axis", "offset", "qidx", "g_frame", "g_shape", "g_dim", "g_mesh",
                    "mesh_vstart", "mesh_tstart", "verts", "tris", "allowed", "mesh_bc", "mesh_br", "tri_c", "tri_b"):
            setattr(s, key, _p(k[key]))
        if not self.culls:
            s.mesh_bc = s.mesh_br = s.tri_c = s.tri_b = None
        self.c = s

    @property
    def allowed(self) -> np.ndarray:
        return self._keep["allowed"]

    def allow(self, i: int, j: int):
        self._keep["allowed"][i, j] = self._keep["allowed"][j, i] = 1

    def frame_id(self, name: str) -> int:
        return self.names.index(name)

    def config_ids(self, names: Sequence[str]) -> np.ndarray:
        return np.asarray([self.config_names.index(n) for n in names], dtype=np.int32)

    # -- FK (aa_rx_sg_tf) ---------------------------------------------------------
    def fk(self, q: np.ndarray):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == (self.n_q,)
        rel = np.empty((self.n_f, 7))
        ab = np.empty((self.n_f, 7))
        self.L.orc_fk(C.byref(self.c), _p(q), _p(rel), _p(ab))
        return rel, ab

    # -- aa_rx_cl_check -----------------------------------------------------------
    def check(self, tf_abs: np.ndarray, tol: float = TOL, want_set: bool = True, early_out: bool = False):
        tf_abs = np.ascontiguousarray(tf_abs, dtype=np.float64)
        ps = np.zeros((self.n_f, self.n_f), dtype=np.uint8) if want_set else None
        ex = C.c_int(0)
        st = self.L.orc_check(C.byref(self.c), _p(tf_abs), tol, None, _p(ps) if want_set else None,
                              int(early_out), C.byref(ex))
        return st, bool(ex.value), ps

    def allow_config(self, q: np.ndarray, tol: float = TOL):
        """aa_rx_sg_allow_config (demo/tmpexec.c:118): allow every frame pair that collides
        (exact verdict or contact band) at q."""
        _, ab = self.fk(q)
        _, _, ps = self.check(ab, tol)
        self._keep["allowed"][:] = np.maximum(self._keep["allowed"], (ps != 0).astype(np.uint8))
        return ps

    def check_batch(self, sub_ids, q_base, Q, tol: float = TOL, early_out=False, hoist=True, threads=1):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = np.ascontiguousarray(q_base, dtype=np.float64)
        n = Q.shape[0]
        st = np.zeros(n, dtype=np.uint8)
        ex = np.zeros(n, dtype=np.uint8)
        self.L.orc_check_batch(C.byref(self.c), n, len(sub), _p(sub), _p(qb), _p(Q), Q.shape[1], tol,
                               int(early_out), int(hoist), threads, _p(st), _p(ex))
        return st, ex

    def check_edges(self, sub_ids, q_base, Q0, Q1, seg_len, tol: float = TOL, early_out=False, hoist=True, threads=1):
        Q0 = np.ascontiguousarray(Q0, dtype=np.float64)
        Q1 = np.ascontiguousarray(Q1, dtype=np.float64)
        assert Q0.shape == Q1.shape
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = np.ascontiguousarray(q_base, dtype=np.float64)
        n = Q0.shape[0]
        st = np.zeros(n, dtype=np.uint8)
        ex = np.zeros(n, dtype=np.uint8)
        first = np.zeros(n, dtype=np.int32)
        nchk = np.zeros(n, dtype=np.int64)
        self.L.orc_check_edges(C.byref(self.c), n, len(sub), _p(sub), _p(qb), _p(Q0), _p(Q1), Q0.shape[1], seg_len,
                               tol, int(early_out), int(hoist), threads, _p(st), _p(ex), _p(first), _p(nchk))
        return st, ex, first, nchk


def pair(sa, da, ta, sb, db, tb, tol: float = TOL):
    """Tri-state + exact verdict of one primitive pair."""
    L = lib()
    da = np.ascontiguousarray(da, dtype=np.float64)
    db = np.ascontiguousarray(db, dtype=np.float64)
    ta = np.ascontiguousarray(ta if ta is not None else np.zeros(7), dtype=np.float64)
    tb = np.ascontiguousarray(tb if tb is not None else np.zeros(7), dtype=np.float64)
    ex = C.c_int(0)
    st = L.orc_pair(sa, _p(da), _p(ta), sb, _p(db), _p(tb), tol, C.byref(ex))
    return st, bool(ex.value)


def sat_box_box(da, ta, db, tb) -> float:
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (da, ta, db, tb)]
    return lib().orc_sat_box_box(*[_p(x) for x in a])


def point_dist(shape, dim, tf, p) -> float:
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dim, tf, p)]
    return lib().orc_point_dist(shape, *[_p(x) for x in a])


def ompl_order(nd: int) -> np.ndarray:
    out = np.zeros(max(nd, 1), dtype=np.int32)
    n = lib().orc_ompl_order(nd, _p(out))
    return out[:n]
