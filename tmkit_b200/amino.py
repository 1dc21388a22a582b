"""ctypes binding of libtmkcl.so - the Python mirror of the Amino-shaped C ABI
(include/tmkit_b200/amino_rx.h).  Names and argument meaning follow the C entry points
one to one (which follow the reference call sites in demo/tmpexec.c:94-150,362-428), so the
parity tests read like the reference's own executor.

This module is plumbing: it loads the shared library and marshals numpy arrays.  There is
no Python (or any CPU) implementation of the path behind it; if the library or a CUDA
device is missing the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence

import numpy as np

from . import scenes as S

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TMK_LIB") or os.path.join(HERE, "libtmkcl.so")  # TMK_LIB: A/B builds while tuning

Q_ROWS_F64, Q_ROWS_F32, Q_SOA_F32 = 0, 1, 2
TF_ROWS_F64, TF_SOA_F32 = 0, 1

_vp, _sz, _i, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
_cs = C.c_char_p

# name -> (restype, argtypes); every symbol include/tmkit_b200/amino_rx.h declares
SIGNATURES = {
    "aa_rx_sg_create": (_vp, []),
    "aa_rx_sg_destroy": (None, [_vp]),
    "aa_rx_sg_add_frame_fixed": (None, [_vp, _cs, _cs, _vp, _vp]),
    "aa_rx_sg_add_frame_revolute": (None, [_vp, _cs, _cs, _vp, _vp, _cs, _vp, _d]),
    "aa_rx_sg_add_frame_prismatic": (None, [_vp, _cs, _cs, _vp, _vp, _cs, _vp, _d]),
    "aa_rx_sg_set_limit_pos": (None, [_vp, _cs, _d, _d]),
    "aa_rx_sg_get_limit_pos": (_i, [_vp, _i, _vp, _vp]),
    "aa_rx_geom_box": (_vp, [_vp, _vp]),
    "aa_rx_geom_sphere": (_vp, [_vp, _d]),
    "aa_rx_geom_cylinder": (_vp, [_vp, _d, _d]),
    "aa_rx_mesh_create": (_vp, []),
    "aa_rx_mesh_destroy": (None, [_vp]),
    "aa_rx_mesh_set_vertices": (None, [_vp, _sz, _vp, _i]),
    "aa_rx_mesh_set_indices": (None, [_vp, _sz, _vp, _i]),
    "aa_rx_geom_mesh": (_vp, [_vp, _vp]),
    "aa_rx_geom_attach": (None, [_vp, _cs, _vp]),
    "aa_rx_sg_allow_collision_name": (None, [_vp, _cs, _cs, _i]),
    "aa_rx_sg_allow_collision": (None, [_vp, _i, _i, _i]),
    "aa_rx_dl_sg": (_vp, [_cs, _cs, _vp]),
    "aa_rx_sg_init": (None, [_vp]),
    "aa_rx_sg_cl_init": (None, [_vp]),
    "aa_rx_sg_config_count": (_sz, [_vp]),
    "aa_rx_sg_frame_count": (_sz, [_vp]),
    "aa_rx_sg_frame_id": (_i, [_vp, _cs]),
    "aa_rx_sg_frame_name": (_cs, [_vp, _i]),
    "aa_rx_sg_frame_parent": (_i, [_vp, _i]),
    "aa_rx_sg_config_id": (_i, [_vp, _cs]),
    "aa_rx_sg_config_name": (_cs, [_vp, _i]),
    "aa_rx_sg_config_indices": (None, [_vp, _sz, _vp, _vp]),
    "aa_rx_sg_config_names": (None, [_vp, _sz, _vp]),
    "aa_rx_sg_config_set": (None, [_vp, _sz, _sz, _vp, _vp, _vp]),
    "aa_rx_sg_tf": (None, [_vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "aa_rx_sg_rel_tf": (None, [_vp, _i, _i, _vp, _sz, _vp]),
    "aa_rx_sg_copy": (_vp, [_vp]),
    "aa_rx_sg_reparent_name": (None, [_vp, _cs, _cs, _vp]),
    "aa_rx_sg_allow_config": (None, [_vp, _sz, _vp]),
    "aa_rx_cl_create": (_vp, [_vp]),
    "aa_rx_cl_destroy": (None, [_vp]),
    "aa_rx_cl_allow": (None, [_vp, _i, _i, _i]),
    "aa_rx_cl_allow_set": (None, [_vp, _vp]),
    "aa_rx_cl_allow_name": (None, [_vp, _cs, _cs, _i]),
    "aa_rx_cl_check": (_i, [_vp, _sz, _vp, _sz, _vp]),
    "aa_rx_cl_set_create": (_vp, [_vp]),
    "aa_rx_cl_set_destroy": (None, [_vp]),
    "aa_rx_cl_set_clear": (None, [_vp]),
    "aa_rx_cl_set_get": (_i, [_vp, _i, _i]),
    "aa_rx_cl_set_set": (None, [_vp, _i, _i, _i]),
    "aa_rx_cl_set_fill": (None, [_vp, _vp]),
    "aa_rx_sg_tf_batch": (None, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _sz, _vp, _vp]),
    "aa_rx_sg_tf_batch_dev": (None, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _i, _sz, _sz, _vp, _vp, _i, _sz, _vp]),
    "aa_rx_cl_check_batch": (_sz, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "aa_rx_cl_check_batch_f32": (_sz, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "aa_rx_cl_check_edges": (_sz, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _d, _vp, _vp]),
    "aa_rx_cl_check_batch_dev": (None, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _i, _sz, _vp, _vp]),
    "aa_rx_cl_check_edges_dev": (None, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _i, _sz, _d, _vp, _vp, _vp]),
    "aa_rx_cl_sync": (None, [_vp, _vp]),
    "aa_rx_cl_track_collisions": (None, [_vp, _i]),
    "aa_rx_cl_batch_collisions": (None, [_vp, _vp]),
    "aa_rx_cl_batch_stats": (None, [_vp, _vp]),
    "aa_rx_cl_stats_enable": (None, [_vp, _i]),
    "aa_rx_sg_frame_type": (_i, [_vp, _i]),
    "aa_rx_sg_frame_config": (_i, [_vp, _i]),
    "aa_rx_sg_frame_axis": (None, [_vp, _i, _vp]),
    "aa_rx_sg_chain_create": (_vp, [_vp, _i, _i]),
    "aa_rx_sg_sub_destroy": (None, [_vp]),
    "aa_rx_sg_sub_config_count": (_sz, [_vp]),
    "aa_rx_sg_sub_configs": (_vp, [_vp]),
    "aa_rx_sg_sub_tip": (_i, [_vp]),
    "aa_rx_mp_create": (_vp, [_vp]),
    "aa_rx_mp_destroy": (None, [_vp]),
    "aa_rx_mp_set_start": (None, [_vp, _sz, _vp]),
    "aa_rx_mp_set_wsgoal": (_i, [_vp, _vp]),
    "aa_rx_mp_set_goal": (_i, [_vp, _sz, _vp]),
    "aa_rx_mp_set_simplify": (None, [_vp, _i]),
    "aa_rx_mp_set_track_collisions": (None, [_vp, _i]),
    "aa_rx_mp_set_seed": (None, [_vp, C.c_uint64]),
    "aa_rx_mp_plan": (_i, [_vp, _d, _vp, _vp]),
    "aa_rx_mp_collisions": (None, [_vp, _vp]),
    "aa_rx_mp_stats": (None, [_vp, _vp]),
    "aa_rx_mp_edge_log": (None, [_vp, _i]),
    "aa_rx_mp_edge_log_get": (_sz, [_vp, _vp, _vp]),
    "tmk_plan_create": (_vp, []),
    "tmk_plan_destroy": (None, [_vp]),
    "tmk_plan_add_action": (None, [_vp, _cs]),
    "tmk_plan_add_reparent": (None, [_vp, _cs, _cs]),
    "tmk_plan_add_motion": (None, [_vp, _sz, _vp, _sz, _vp, _sz]),
    "tmk_plan_write": (_i, [_vp, _cs]),
    "tmk_plan_parse": (_vp, [_cs]),
    "tmk_plan_op_count": (_sz, [_vp]),
    "tmk_plan_op_kind": (_i, [_vp, _sz]),
    "tmk_plan_op_text": (_cs, [_vp, _sz, _sz]),
    "tmk_plan_motion_names": (_sz, [_vp, _sz]),
    "tmk_plan_motion_points": (_sz, [_vp, _sz]),
    "tmk_plan_motion_data": (_vp, [_vp, _sz]),
    "tmk_shard_per_rank": (_sz, [_sz, _i]),
    "tmk_shard_bounds": (None, [_sz, _i, _i, _vp, _vp]),
    "tmk_shard_words": (_sz, [_sz, _i]),
    "aa_rx_cl_multi_create": (_vp, [_vp, _i, _vp]),
    "aa_rx_cl_multi_destroy": (None, [_vp]),
    "aa_rx_cl_multi_device_count": (_i, [_vp]),
    "aa_rx_cl_multi_device": (_i, [_vp, _i]),
    "aa_rx_cl_multi_context": (_vp, [_vp, _i]),
    "aa_rx_cl_multi_allow": (None, [_vp, _i, _i, _i]),
    "aa_rx_cl_multi_allow_set": (None, [_vp, _vp]),
    "aa_rx_cl_multi_allow_name": (None, [_vp, _cs, _cs, _i]),
    "aa_rx_cl_multi_track_collisions": (None, [_vp, _i]),
    "aa_rx_cl_multi_batch_collisions": (None, [_vp, _vp]),
    "aa_rx_cl_check_batch_multi": (_sz, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "aa_rx_cl_check_batch_multi_f32": (_sz, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "aa_rx_cl_check_edges_multi": (_sz, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _d, _vp, _vp]),
    "aa_rx_cl_check_batch_multi_dev": (None, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _i, _sz, _vp]),
    "aa_rx_cl_check_edges_multi_dev": (None, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _i, _sz, _d, _vp, _vp]),
    "aa_rx_cl_multi_sync": (None, [_vp]),
    "tmk_shard_unique_id": (None, [_vp]),
    "tmk_shard_comm_create": (_vp, [_i, _i, _vp]),
    "tmk_shard_comm_destroy": (None, [_vp]),
    "tmk_shard_comm_rank": (_i, [_vp]),
    "tmk_shard_comm_world": (_i, [_vp]),
    "tmk_shard_comm_overlap": (None, [_vp, _i]),
    "tmk_shard_comm_join": (None, [_vp, _vp]),
    "aa_rx_cl_check_batch_shard_dev": (None, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp, _i, _sz, _vp, _vp]),
    "aa_rx_cl_check_edges_shard_dev": (None, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _i, _sz, _d, _vp, _vp, _vp]),
    "aa_rx_host_alloc": (_vp, [_sz]),
    "aa_rx_host_free": (None, [_vp]),
    "aa_rx_host_register": (_i, [_vp, _sz]),
    "aa_rx_host_unregister": (None, [_vp]),
    "aa_rx_cl_kernel_timing": (None, [_vp, _i]),
    "aa_rx_cl_kernel_time": (_sz, [_vp, _vp, _vp]),
    "tmk_fp32_peak_tflops": (_d, []),
    "tmk_version": (_cs, []),
}


class BatchStats(C.Structure):
    _fields_ = [
        ("n_units", C.c_uint64), ("n_states", C.c_uint64), ("n_joints", C.c_uint64),
        ("n_moving_geoms", C.c_uint64), ("n_pairs", C.c_uint64), ("n_narrow", C.c_uint64 * 8),
        ("n_uncertain", C.c_uint64), ("n_invalid", C.c_uint64), ("n_launches", C.c_uint64),
        ("gpu_ms", C.c_double),
    ]


_lib = None
SHARD_ALIGN = 1024
SHARD_ID_BYTES = 128


def _preload_nccl():
    """libtmkcl.so needs libnccl.so.2 (the all-gather of the validity words).  A process that also imports torch
    must end up with ONE NCCL: torch's bundled copy is loaded first when it exists, so that the soname is already
    resolved (to the newer library) whichever of the two is imported first; otherwise the system one is used."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return cand
    except Exception:
        pass
    return None


def lib():
    """Load libtmkcl.so (built by __graft_entry__.build() / tmkit_b200/csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it (python -c 'import __graft_entry__ as g; g.build()'). "
                               "There is no CPU fallback.")
        _preload_nccl()
        # RTLD_GLOBAL: dlopen'ed scene plugins resolve the aa_rx_* constructors against this library
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _b(s: Optional[str]) -> bytes:
    return (s or "").encode()


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class CollisionSet:
    """struct aa_rx_cl_set (demo/tmpexec.c:399-415)."""

    def __init__(self, sg: "SceneGraph"):
        self.L = lib()
        self.sg = sg
        self.h = self.L.aa_rx_cl_set_create(sg.h)

    def clear(self):
        self.L.aa_rx_cl_set_clear(self.h)

    def get(self, i: int, j: int) -> int:
        return self.L.aa_rx_cl_set_get(self.h, i, j)

    def set(self, i: int, j: int, v: int = 1):
        self.L.aa_rx_cl_set_set(self.h, i, j, v)

    def fill(self, other: "CollisionSet"):
        self.L.aa_rx_cl_set_fill(self.h, other.h)

    def pairs(self):
        n = self.sg.frame_count()
        return [(i, j) for i in range(n) for j in range(i) if self.get(i, j)]

    def matrix(self) -> np.ndarray:
        n = self.sg.frame_count()
        m = np.zeros((n, n), dtype=np.uint8)
        for i, j in self.pairs():
            m[i, j] = m[j, i] = 1
        return m

    def __del__(self):
        try:
            if self.h:
                self.L.aa_rx_cl_set_destroy(self.h)
                self.h = None
        except Exception:
            pass


class SceneGraph:
    """struct aa_rx_sg."""

    def __init__(self, handle=None):
        self.L = lib()
        self.h = handle if handle is not None else self.L.aa_rx_sg_create()
        self._own = True

    # -- B2: construction --------------------------------------------------------
    @classmethod
    def from_scene(cls, sc: S.Scene, init: bool = True) -> "SceneGraph":
        """Push a neutral scene description through the constructor ABI, exactly as a
        generated scene plugin would."""
        sg = cls()
        L = sg.L
        for f in sc.frames:
            q = _f64(f.quat)
            v = _f64(f.xyz)
            if f.type == S.FIXED:
                L.aa_rx_sg_add_frame_fixed(sg.h, _b(f.parent), _b(f.name), _p(q), _p(v))
            else:
                ax = _f64(f.axis)
                fn = L.aa_rx_sg_add_frame_revolute if f.type == S.REVOLUTE else L.aa_rx_sg_add_frame_prismatic
                fn(sg.h, _b(f.parent), _b(f.name), _p(q), _p(v), _b(f.config), _p(ax), float(f.offset))
                if f.limits is not None:
                    L.aa_rx_sg_set_limit_pos(sg.h, _b(f.config), float(f.limits[0]), float(f.limits[1]))
            for g in f.geoms:
                if g.shape == S.BOX:
                    d = _f64(g.dims)
                    gh = L.aa_rx_geom_box(None, _p(d))
                elif g.shape == S.SPHERE:
                    gh = L.aa_rx_geom_sphere(None, float(g.dims[0]))
                elif g.shape == S.CYLINDER:
                    gh = L.aa_rx_geom_cylinder(None, float(g.dims[0]), float(g.dims[1]))
                else:
                    m = L.aa_rx_mesh_create()
                    vv = np.ascontiguousarray(g.verts, dtype=np.float32)
                    ii = np.ascontiguousarray(g.tris, dtype=np.uint32)
                    L.aa_rx_mesh_set_vertices(m, len(vv), _p(vv), 1)
                    L.aa_rx_mesh_set_indices(m, len(ii), _p(ii), 1)
                    gh = L.aa_rx_geom_mesh(None, m)
                    L.aa_rx_mesh_destroy(m)  # the geometry holds its own reference
                L.aa_rx_geom_attach(sg.h, _b(f.name), gh)
        for a, b in sc.allowed:
            L.aa_rx_sg_allow_collision_name(sg.h, _b(a), _b(b), 1)
        if init:
            sg.init()
            sg.cl_init()
        return sg

    @classmethod
    def dl(cls, plugin_file: str, scene_name: str, sg: Optional["SceneGraph"] = None) -> Optional["SceneGraph"]:
        L = lib()
        h = L.aa_rx_dl_sg(_b(plugin_file), _b(scene_name), sg.h if sg else None)
        if not h:
            return None
        if sg:
            return sg
        return cls(handle=h)

    def init(self):
        self.L.aa_rx_sg_init(self.h)

    def cl_init(self):
        self.L.aa_rx_sg_cl_init(self.h)

    # -- queries -----------------------------------------------------------------
    def config_count(self) -> int:
        return self.L.aa_rx_sg_config_count(self.h)

    def frame_count(self) -> int:
        return self.L.aa_rx_sg_frame_count(self.h)

    def frame_id(self, name: str) -> int:
        return self.L.aa_rx_sg_frame_id(self.h, _b(name))

    def frame_name(self, i: int) -> str:
        return self.L.aa_rx_sg_frame_name(self.h, i).decode()

    def frame_parent(self, i: int) -> int:
        return self.L.aa_rx_sg_frame_parent(self.h, i)

    def frame_names(self):
        return [self.frame_name(i) for i in range(self.frame_count())]

    def config_id(self, name: str) -> int:
        return self.L.aa_rx_sg_config_id(self.h, _b(name))

    def config_name(self, i: int) -> str:
        return self.L.aa_rx_sg_config_name(self.h, i).decode()

    def config_names(self):
        return [self.config_name(i) for i in range(self.config_count())]

    def config_indices(self, names: Sequence[str]) -> np.ndarray:
        arr = (C.c_char_p * len(names))(*[_b(n) for n in names])
        ids = np.zeros(len(names), dtype=np.int32)
        self.L.aa_rx_sg_config_indices(self.h, len(names), arr, _p(ids))
        return ids

    def config_set(self, ids: Sequence[int], q_sub, q_all) -> np.ndarray:
        q_all = _f64(q_all).copy()
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        q_sub = _f64(q_sub)
        self.L.aa_rx_sg_config_set(self.h, len(q_all), len(ids), _p(ids), _p(q_sub), _p(q_all))
        return q_all

    def limits(self, ids: Iterable[int]) -> np.ndarray:
        out = []
        for i in ids:
            lo, hi = C.c_double(), C.c_double()
            if self.L.aa_rx_sg_get_limit_pos(self.h, int(i), C.byref(lo), C.byref(hi)) != 0:
                out.append((-np.pi, np.pi))
            else:
                out.append((lo.value, hi.value))
        return np.asarray(out)

    # -- FK ------------------------------------------------------------------------
    def tf(self, q):
        """aa_rx_sg_tf(sg, n_q, q, n_tf, TF_rel, 7, TF_abs, 7)  (demo/tmpexec.c:322-325)."""
        q = _f64(q)
        n_f = self.frame_count()
        rel = np.zeros((n_f, 7))
        ab = np.zeros((n_f, 7))
        self.L.aa_rx_sg_tf(self.h, len(q), _p(q), n_f, _p(rel), 7, _p(ab), 7)
        return rel, ab

    def rel_tf(self, parent: int, child: int, tf_abs) -> np.ndarray:
        tf_abs = _f64(tf_abs)
        out = np.zeros(7)
        self.L.aa_rx_sg_rel_tf(self.h, parent, child, _p(tf_abs), tf_abs.shape[1], _p(out))
        return out

    def tf_batch(self, sub_ids, q_base, Q, frames=None) -> np.ndarray:
        Q = _f64(Q)
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        n_out = self.frame_count() if frames is None else len(frames)
        fr = None if frames is None else np.ascontiguousarray(frames, dtype=np.int32)
        out = np.zeros((Q.shape[0], n_out, 7))
        self.L.aa_rx_sg_tf_batch(self.h, Q.shape[0], len(sub), _p(sub), len(qb), _p(qb), _p(Q), Q.shape[1],
                                 n_out, _p(fr), _p(out))
        return out

    def tf_batch_dev(self, n, sub_ids, q_base, dQ: int, layout: int, ld: int, frames, dTF: int, tf_layout: int, ld_tf: int,
                     stream: int = 0):
        """aa_rx_sg_tf_batch_dev: raw device addresses in and out, asynchronous on `stream`."""
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        fr = None if frames is None else np.ascontiguousarray(frames, dtype=np.int32)
        self.L.aa_rx_sg_tf_batch_dev(self.h, n, len(sub), _p(sub), len(qb), _p(qb), dQ, layout, ld,
                                     0 if fr is None else len(fr), _p(fr), dTF, tf_layout, ld_tf, stream or None)

    # -- mutation ------------------------------------------------------------------
    def copy(self) -> "SceneGraph":
        return SceneGraph(handle=self.L.aa_rx_sg_copy(self.h))

    def reparent_name(self, new_parent: str, child: str, tf7):
        tf7 = _f64(tf7)
        self.L.aa_rx_sg_reparent_name(self.h, _b(new_parent), _b(child), _p(tf7))

    def allow_collision_name(self, a: str, b: str, allowed: int = 1):
        self.L.aa_rx_sg_allow_collision_name(self.h, _b(a), _b(b), allowed)

    def allow_config(self, q):
        q = _f64(q)
        self.L.aa_rx_sg_allow_config(self.h, len(q), _p(q))

    def __del__(self):
        try:
            if self.h and self._own:
                self.L.aa_rx_sg_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Collision:
    """struct aa_rx_cl."""

    def __init__(self, sg: SceneGraph):
        self.L = lib()
        self.sg = sg
        self.h = self.L.aa_rx_cl_create(sg.h)

    def allow(self, i: int, j: int, allowed: int = 1):
        self.L.aa_rx_cl_allow(self.h, i, j, allowed)

    def allow_set(self, s: CollisionSet):
        self.L.aa_rx_cl_allow_set(self.h, s.h)

    def allow_name(self, a: str, b: str, allowed: int = 1):
        self.L.aa_rx_cl_allow_name(self.h, _b(a), _b(b), allowed)

    def check(self, tf_abs, cl_set: Optional[CollisionSet] = None) -> int:
        """aa_rx_cl_check(cl, n_f, TF_abs, 7, cl_set)  (demo/tmpexec.c:409)."""
        tf_abs = _f64(tf_abs)
        return self.L.aa_rx_cl_check(self.h, tf_abs.shape[0], _p(tf_abs), tf_abs.shape[1],
                                     cl_set.h if cl_set else None)

    @staticmethod
    def unpack_bits(bits: np.ndarray, n: int) -> np.ndarray:
        return np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)

    def check_batch(self, sub_ids, q_base, Q):
        """Validity bitmap (True = collision-free) of the rows of Q."""
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        Q = np.ascontiguousarray(Q)
        n = Q.shape[0]
        bits = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        if Q.dtype == np.float32:
            bad = self.L.aa_rx_cl_check_batch_f32(self.h, n, len(sub), _p(sub), len(qb), _p(qb), _p(Q), Q.shape[1], _p(bits))
        else:
            Q = _f64(Q)
            bad = self.L.aa_rx_cl_check_batch(self.h, n, len(sub), _p(sub), len(qb), _p(qb), _p(Q), Q.shape[1], _p(bits))
        valid = self.unpack_bits(bits, n)
        assert int((~valid).sum()) == bad
        return valid

    def check_edges(self, sub_ids, q_base, Q0, Q1, seg_len: float, want_first: bool = True):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        Q0, Q1 = _f64(Q0), _f64(Q1)
        n = Q0.shape[0]
        bits = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        first = np.full(n + 1, -7, dtype=np.int32) if want_first else None
        self.L.aa_rx_cl_check_edges(self.h, n, len(sub), _p(sub), len(qb), _p(qb), _p(Q0), _p(Q1), Q0.shape[1],
                                    float(seg_len), _p(bits), _p(first))
        return self.unpack_bits(bits, n), (first[:n] if want_first else None)

    # device-resident variants: pointers are raw device addresses (e.g. torch tensor .data_ptr())
    def check_batch_dev(self, n, sub_ids, q_base, dQ: int, layout: int, ld: int, d_bits: int, stream: int = 0):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        self.L.aa_rx_cl_check_batch_dev(self.h, n, len(sub), _p(sub), len(qb), _p(qb), dQ, layout, ld, d_bits,
                                        stream or None)

    def check_edges_dev(self, n, sub_ids, q_base, dQ0: int, dQ1: int, layout: int, ld: int, seg_len: float,
                        d_bits: int, d_first: int = 0, stream: int = 0):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        self.L.aa_rx_cl_check_edges_dev(self.h, n, len(sub), _p(sub), len(qb), _p(qb), dQ0, dQ1, layout, ld,
                                        float(seg_len), d_bits, d_first or None, stream or None)

    def sync(self, stream: int = 0):
        self.L.aa_rx_cl_sync(self.h, stream or None)

    def track_collisions(self, enable: bool = True):
        self.L.aa_rx_cl_track_collisions(self.h, int(enable))

    def batch_collisions(self) -> CollisionSet:
        s = CollisionSet(self.sg)
        self.L.aa_rx_cl_batch_collisions(self.h, s.h)
        return s

    def kernel_timing(self, enable: bool = True):
        self.L.aa_rx_cl_kernel_timing(self.h, int(enable))

    def kernel_time(self):
        """(timed batches, ms in the main kernel, ms in the re-check kernel) since the last call."""
        a, b = C.c_double(0), C.c_double(0)
        n = self.L.aa_rx_cl_kernel_time(self.h, C.byref(a), C.byref(b))
        return n, a.value, b.value

    def stats_enable(self, enable: bool = True):
        self.L.aa_rx_cl_stats_enable(self.h, int(enable))

    def stats(self) -> BatchStats:
        st = BatchStats()
        self.L.aa_rx_cl_batch_stats(self.h, C.byref(st))
        return st

    def __del__(self):
        try:
            if self.h:
                self.L.aa_rx_cl_destroy(self.h)
                self.h = None
        except Exception:
            pass


def shard_bounds(n: int, world: int, rank: int):
    """tmk_shard_bounds: the slice [lo, hi) of rank `rank` (no device needed)."""
    lo, hi = C.c_size_t(0), C.c_size_t(0)
    lib().tmk_shard_bounds(n, world, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def shard_words(n: int, world: int) -> int:
    return lib().tmk_shard_words(n, world)


class CollisionMulti:
    """struct aa_rx_cl_multi: one handle over several GPUs of the box (one process)."""

    def __init__(self, sg: SceneGraph, devices: Optional[Sequence[int]] = None):
        self.L = lib()
        self.sg = sg
        if devices is None:
            self.h = self.L.aa_rx_cl_multi_create(sg.h, 0, None)
        else:
            d = np.ascontiguousarray(devices, dtype=np.int32)
            self.h = self.L.aa_rx_cl_multi_create(sg.h, len(d), _p(d))
        self.n_dev = self.L.aa_rx_cl_multi_device_count(self.h)
        self.devices = [self.L.aa_rx_cl_multi_device(self.h, k) for k in range(self.n_dev)]

    def allow_set(self, s: CollisionSet):
        self.L.aa_rx_cl_multi_allow_set(self.h, s.h)

    def allow_name(self, a: str, b: str, allowed: int = 1):
        self.L.aa_rx_cl_multi_allow_name(self.h, _b(a), _b(b), allowed)

    def track_collisions(self, enable: bool = True):
        self.L.aa_rx_cl_multi_track_collisions(self.h, int(enable))

    def batch_collisions(self) -> CollisionSet:
        s = CollisionSet(self.sg)
        self.L.aa_rx_cl_multi_batch_collisions(self.h, s.h)
        return s

    def check_batch(self, sub_ids, q_base, Q):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        Q = np.ascontiguousarray(Q)
        n = Q.shape[0]
        bits = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        bits[-1] = 0xdeadbeef  # canary: nothing may be written past ceil(n / 32) words
        if Q.dtype == np.float32:
            bad = self.L.aa_rx_cl_check_batch_multi_f32(self.h, n, len(sub), _p(sub), len(qb), _p(qb), _p(Q), Q.shape[1], _p(bits))
        else:
            Q = _f64(Q)
            bad = self.L.aa_rx_cl_check_batch_multi(self.h, n, len(sub), _p(sub), len(qb), _p(qb), _p(Q), Q.shape[1], _p(bits))
        assert bits[-1] == 0xdeadbeef
        valid = Collision.unpack_bits(bits, n)
        assert int((~valid).sum()) == bad
        return valid

    def check_edges(self, sub_ids, q_base, Q0, Q1, seg_len: float, want_first: bool = True):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        Q0, Q1 = _f64(Q0), _f64(Q1)
        n = Q0.shape[0]
        bits = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        first = np.full(n + 1, -7, dtype=np.int32) if want_first else None
        self.L.aa_rx_cl_check_edges_multi(self.h, n, len(sub), _p(sub), len(qb), _p(qb), _p(Q0), _p(Q1), Q0.shape[1],
                                          float(seg_len), _p(bits), _p(first))
        return Collision.unpack_bits(bits, n), (first[:n] if want_first else None)

    def check_batch_dev(self, n, sub_ids, q_base, dQ: Sequence[int], layout: int, ld: int, d_bits_all: Sequence[int]):
        """dQ[k] / d_bits_all[k]: raw device addresses on device k (slice rows / gathered bitmap)."""
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        a = (C.c_void_p * self.n_dev)(*dQ)
        b = (C.c_void_p * self.n_dev)(*d_bits_all)
        self.L.aa_rx_cl_check_batch_multi_dev(self.h, n, len(sub), _p(sub), len(qb), _p(qb), a, layout, ld, b)

    def check_edges_dev(self, n, sub_ids, q_base, dQ0, dQ1, layout: int, ld: int, seg_len: float, d_bits_all,
                        d_first=None):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        a = (C.c_void_p * self.n_dev)(*dQ0)
        a1 = (C.c_void_p * self.n_dev)(*dQ1)
        b = (C.c_void_p * self.n_dev)(*d_bits_all)
        f = (C.c_void_p * self.n_dev)(*d_first) if d_first is not None else None
        self.L.aa_rx_cl_check_edges_multi_dev(self.h, n, len(sub), _p(sub), len(qb), _p(qb), a, a1, layout, ld,
                                              float(seg_len), b, f)

    def sync(self):
        self.L.aa_rx_cl_multi_sync(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.aa_rx_cl_multi_destroy(self.h)
                self.h = None
        except Exception:
            pass


class ShardComm:
    """struct tmk_shard_comm: one process per GPU; the 128-byte id is made by rank 0 (ShardComm.unique_id) and
    distributed by the caller (torch.distributed broadcast, MPI, a file ...)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(SHARD_ID_BYTES)
        lib().tmk_shard_unique_id(buf)
        return buf.raw

    def __init__(self, rank: int, world: int, id_bytes: bytes):
        self.L = lib()
        assert len(id_bytes) == SHARD_ID_BYTES
        self._id = C.create_string_buffer(id_bytes, SHARD_ID_BYTES)
        self.rank, self.world = rank, world
        self.h = self.L.tmk_shard_comm_create(rank, world, self._id)

    def overlap(self, enable: bool = True):
        self.L.tmk_shard_comm_overlap(self.h, int(enable))

    def join(self, stream: int = 0):
        self.L.tmk_shard_comm_join(self.h, stream or None)

    def check_batch_dev(self, cl: Collision, n_total, sub_ids, q_base, dQ_slice: int, layout: int, ld: int,
                        d_bits_all: int, stream: int = 0):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        self.L.aa_rx_cl_check_batch_shard_dev(cl.h, self.h, n_total, len(sub), _p(sub), len(qb), _p(qb), dQ_slice,
                                              layout, ld, d_bits_all, stream or None)

    def check_edges_dev(self, cl: Collision, n_total, sub_ids, q_base, dQ0: int, dQ1: int, layout: int, ld: int,
                        seg_len: float, d_bits_all: int, d_first: int = 0, stream: int = 0):
        sub = np.ascontiguousarray(sub_ids, dtype=np.int32)
        qb = _f64(q_base)
        self.L.aa_rx_cl_check_edges_shard_dev(cl.h, self.h, n_total, len(sub), _p(sub), len(qb), _p(qb), dQ0, dQ1,
                                              layout, ld, float(seg_len), d_bits_all, d_first or None, stream or None)

    def destroy(self):
        if self.h:
            self.L.tmk_shard_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class MpStats(C.Structure):
    _fields_ = [("n_iterations", C.c_uint64), ("n_edge_calls", C.c_uint64), ("n_edges_checked", C.c_uint64),
                ("n_nodes", C.c_uint64), ("n_ik_iterations", C.c_uint64), ("n_goals", C.c_uint64),
                ("ik_s", C.c_double), ("rrt_s", C.c_double), ("simplify_s", C.c_double)]


class SubSceneGraph:
    """struct aa_rx_sg_sub: the chain root -> tip (tm-blocks.py:202 aa.scene_chain)."""

    def __init__(self, sg: SceneGraph, root: str, tip: str):
        self.L = lib()
        self.sg = sg
        r = -1 if root == "" else sg.frame_id(root)
        self.h = self.L.aa_rx_sg_chain_create(sg.h, r, sg.frame_id(tip))

    def config_ids(self) -> np.ndarray:
        n = self.L.aa_rx_sg_sub_config_count(self.h)
        p = C.cast(self.L.aa_rx_sg_sub_configs(self.h), C.POINTER(C.c_int))
        return np.asarray([p[i] for i in range(n)], dtype=np.int32)

    def config_names(self):
        return [self.sg.config_name(int(i)) for i in self.config_ids()]

    def __del__(self):
        try:
            if self.h:
                self.L.aa_rx_sg_sub_destroy(self.h)
                self.h = None
        except Exception:
            pass


class MotionPlanner:
    """struct aa_rx_mp: what robray:motion-plan drives (lisp/python/tmsmtpy.lisp:55-60)."""

    def __init__(self, ssg: SubSceneGraph):
        self.L = lib()
        self.ssg = ssg
        self.h = self.L.aa_rx_mp_create(ssg.h)

    def set_start(self, q_all):
        q = _f64(q_all)
        self.L.aa_rx_mp_set_start(self.h, len(q), _p(q))

    def set_wsgoal(self, tf7) -> int:
        e = _f64(tf7)
        return self.L.aa_rx_mp_set_wsgoal(self.h, _p(e))

    def set_goal(self, q_sub) -> int:
        q = _f64(q_sub)
        return self.L.aa_rx_mp_set_goal(self.h, len(q), _p(q))

    def set_simplify(self, on: bool = True):
        self.L.aa_rx_mp_set_simplify(self.h, int(on))

    def set_track_collisions(self, on: bool = True):
        self.L.aa_rx_mp_set_track_collisions(self.h, int(on))

    def set_seed(self, seed: int):
        self.L.aa_rx_mp_set_seed(self.h, seed)

    def plan(self, timeout: float = 3.0):
        """returns the path (n_points x n_all) or None on planning failure."""
        n = C.c_size_t(0)
        ptr = C.c_void_p(0)
        rc = self.L.aa_rx_mp_plan(self.h, float(timeout), C.byref(n), C.byref(ptr))
        if rc != 0:
            return None
        n_all = self.ssg.sg.config_count()
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n.value, n_all)).copy()
        C.CDLL(None).free(ptr)
        return arr

    def collisions(self) -> CollisionSet:
        s = CollisionSet(self.ssg.sg)
        self.L.aa_rx_mp_collisions(self.h, s.h)
        return s

    def edge_log(self, enable: bool = True):
        self.L.aa_rx_mp_edge_log(self.h, int(enable))

    def edge_log_get(self):
        a, b = C.c_void_p(0), C.c_void_p(0)
        n = self.L.aa_rx_mp_edge_log_get(self.h, C.byref(a), C.byref(b))
        ns = len(self.ssg.config_ids())
        if n == 0:
            return np.zeros((0, ns)), np.zeros((0, ns))
        q0 = np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_double)), shape=(n, ns)).copy()
        q1 = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_double)), shape=(n, ns)).copy()
        return q0, q1

    def stats(self) -> MpStats:
        st = MpStats()
        self.L.aa_rx_mp_stats(self.h, C.byref(st))
        return st

    def __del__(self):
        try:
            if self.h:
                self.L.aa_rx_mp_destroy(self.h)
                self.h = None
        except Exception:
            pass


class PlanFile:
    """plan files (doc/md/planfile.md)."""

    def __init__(self, handle=None):
        self.L = lib()
        self.h = handle if handle is not None else self.L.tmk_plan_create()

    @classmethod
    def parse(cls, path: str) -> Optional["PlanFile"]:
        h = lib().tmk_plan_parse(_b(path))
        return cls(handle=h) if h else None

    def add_action(self, text: str):
        self.L.tmk_plan_add_action(self.h, _b(text))

    def add_reparent(self, frame: str, new_parent: str):
        self.L.tmk_plan_add_reparent(self.h, _b(frame), _b(new_parent))

    def add_motion(self, names: Sequence[str], points):
        pts = _f64(points)
        arr = (C.c_char_p * len(names))(*[_b(n) for n in names])
        self.L.tmk_plan_add_motion(self.h, len(names), arr, pts.shape[0], _p(pts), pts.shape[1])

    def write(self, path: str) -> int:
        return self.L.tmk_plan_write(self.h, _b(path))

    def ops(self):
        out = []
        for i in range(self.L.tmk_plan_op_count(self.h)):
            kind = chr(self.L.tmk_plan_op_kind(self.h, i))
            if kind == "a":
                out.append(("a", self.L.tmk_plan_op_text(self.h, i, 0).decode()))
            elif kind == "r":
                out.append(("r", self.L.tmk_plan_op_text(self.h, i, 0).decode(), self.L.tmk_plan_op_text(self.h, i, 1).decode()))
            else:
                nn, npt = self.L.tmk_plan_motion_names(self.h, i), self.L.tmk_plan_motion_points(self.h, i)
                names = [self.L.tmk_plan_op_text(self.h, i, k).decode() for k in range(nn)]
                data = np.zeros((npt, nn))
                if npt:
                    p = C.cast(self.L.tmk_plan_motion_data(self.h, i), C.POINTER(C.c_double))
                    data = np.ctypeslib.as_array(p, shape=(npt, nn)).copy()
                out.append(("m", names, data))
        return out

    def __del__(self):
        try:
            if self.h:
                self.L.tmk_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass
