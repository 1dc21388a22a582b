"""C-ABI checks that need no GPU: the library loads, exports every symbol include/*.h declares,
and the host-side scene-graph logic (ingest, topological sort, name/id maps, allowed pairs,
collision sets, copy / reparent, plugin loading) behaves as the reference's call sites expect
(demo/tmpexec.c:94-150, 269-291, 317-344, 399-419).  No compute entry point is called here.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from tmkit_b200 import amino as A
from tmkit_b200 import scenes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tmkit_b200", "amino_rx.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    names = re.findall(r"\b((?:aa_rx|tmk)_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(built):
    decl = _declared_functions()
    assert len(decl) > 50
    out = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in decl if n not in exported]
    assert not missing, missing
    # and the ctypes mirror binds each of them
    L = A.lib()
    assert sorted(A.SIGNATURES) == decl
    for n in decl:
        assert getattr(L, n) is not None
    assert b"sm_100a" in L.tmk_version()


def test_library_exports_the_reference_plan_file_abi(built):
    """every function include/tmkit_b200/tmplan.h declares (the names of /root/reference/include/tmsmt/tmplan.h)"""
    src = open(os.path.join(ROOT, "include", "tmkit_b200", "tmplan.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decl = sorted(set(re.findall(r"\b(tmplan_[a-z0-9_]+)\s*\(", src)))
    assert len(decl) >= 29, decl
    out = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert not [n for n in decl if n not in exported]
    # the reference header's names, as listed in its public API (tmplan.h:34-247), all appear
    for n in ("tmplan_parse_filename", "tmplan_parse_file", "tmplan_map_ops", "tmplan_op_type", "tmplan_op_action_get",
              "tmplan_op_motion_plan_path", "tmplan_op_motion_plan_config_count", "tmplan_op_motion_plan_path_size",
              "tmplan_op_motion_plan_vars", "tmplan_op_motion_plan_map_var", "tmplan_op_reparent_get_frame",
              "tmplan_op_reparent_get_new_parent", "tmplan_write", "tmplan_create", "tmplan_op_count"):
        assert n in decl, n


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "--list-elf", A.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_does_not_link_the_oracle(built):
    out = subprocess.run(["nm", "-D", A.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "orc_" not in out
    ldd = subprocess.run(["ldd", A.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
    for root, _, files in os.walk(os.path.join(ROOT, "tmkit_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "tmk_oracle" not in txt, f


def test_shard_bounds_cover_the_batch_and_match_the_python_restatement(built):
    """tmk_shard_bounds (multi.cu) = tmkit_b200/shard.py slice_bounds: contiguous, 1024-aligned, covering [0, n)"""
    from tmkit_b200 import shard
    for n in (0, 1, 1023, 1024, 1025, 100_000, (1 << 20) + 7, 1 << 24):
        for world in (1, 2, 3, 4, 8):
            per = A.lib().tmk_shard_per_rank(n, world)
            assert per % A.SHARD_ALIGN == 0 and A.shard_words(n, world) * 32 == per
            assert A.shard_words(n, world) == shard.words_per_rank(n, world)
            want = shard.slice_bounds(n, world)
            prev = 0
            for r in range(world):
                lo, hi = A.shard_bounds(n, world, r)
                assert (lo, hi) == want[r]
                assert lo == prev and lo <= hi <= n and (lo % A.SHARD_ALIGN == 0 or lo == n)
                prev = hi
            assert prev == n


@pytest.fixture(scope="module")
def sg(built):
    return A.SceneGraph.from_scene(S.baxter_sussman())


def test_ingest_counts_and_order(sg):
    flat = S.flatten(S.baxter_sussman())
    assert sg.frame_count() == len(flat["names"]) == 58
    assert sg.config_count() == 15
    assert sg.frame_names() == flat["names"]
    assert sg.config_names() == flat["config_names"]
    for i in range(sg.frame_count()):
        p = sg.frame_parent(i)
        assert p < i  # topologically sorted: parent before child
        assert p == flat["parent"][i]
    assert sg.frame_id("no_such_frame") < 0
    assert sg.frame_id("") == -1  # AA_RX_FRAME_ROOT


def test_config_maps(sg):
    """demo/tmpexec.c:269-291: map by NAME, then scatter q_sub into q_all."""
    names = ["right_w2", "right_w1", "right_w0", "right_s1", "right_s0", "right_e1", "right_e0"]  # q0.tmp order
    ids = sg.config_indices(names)
    assert [sg.config_name(int(i)) for i in ids] == names
    q_all = np.zeros(sg.config_count())
    vals = np.arange(1.0, 8.0)
    out = sg.config_set(ids, vals, q_all)
    for n, v in zip(names, vals):
        assert out[sg.config_id(n)] == v
    assert out.sum() == vals.sum()
    lim = sg.limits(sg.config_indices(S.RIGHT_ARM))
    np.testing.assert_allclose(lim, S.arm_limits())


def test_collision_set_bit_matrix(sg):
    cs = A.CollisionSet(sg)
    n = sg.frame_count()
    assert cs.pairs() == []
    cs.set(3, 40)
    assert cs.get(3, 40) == 1 and cs.get(40, 3) == 1 and cs.get(3, 41) == 0
    cs2 = A.CollisionSet(sg)
    cs2.set(n - 1, 0)
    cs2.fill(cs)
    assert sorted(cs2.pairs()) == [(40, 3), (n - 1, 0)]
    cs2.set(3, 40, 0)
    assert cs2.pairs() == [(n - 1, 0)]
    cs2.clear()
    assert cs2.pairs() == []


def test_rel_tf_is_host_math(sg):
    """aa_rx_sg_rel_tf (tmpexec.c:327-330) works on caller-provided absolute transforms."""
    n = sg.frame_count()
    rng = np.random.default_rng(0)
    ab = np.zeros((n, 7))
    q = rng.normal(size=(n, 4))
    ab[:, :4] = q / np.linalg.norm(q, axis=1, keepdims=True)
    ab[:, 4:] = rng.uniform(-1, 1, (n, 3))
    rel = sg.rel_tf(5, 9, ab)
    # parent * rel == child

    def mul(a, b):
        x1, y1, z1, w1 = a[:4]
        x2, y2, z2, w2 = b[:4]
        r = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                      w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
        u = a[:3]
        t = 2 * np.cross(u, b[4:])
        v = b[4:] + a[3] * t + np.cross(u, t) + a[4:]
        return np.concatenate([r, v])

    np.testing.assert_allclose(mul(ab[5], rel), ab[9], atol=1e-14)
    np.testing.assert_allclose(sg.rel_tf(-1, 9, ab), ab[9], atol=1e-15)


def test_copy_and_reparent_host_side(sg):
    sg2 = sg.copy()
    sg2.reparent_name("end_effector_grasp", "block_c", [0, 0, 0, 2.0, 0, 0, 0.05])  # quaternion gets normalised
    sg2.init()
    sg2.cl_init()
    assert sg2.frame_count() == sg.frame_count()
    assert sg2.frame_parent(sg2.frame_id("block_c")) == sg2.frame_id("end_effector_grasp")
    assert sg.frame_parent(sg.frame_id("block_c")) == sg.frame_id("block_a")  # the original is untouched
    assert sorted(sg2.frame_names()) == sorted(sg.frame_names())
    for i in range(sg2.frame_count()):
        assert sg2.frame_parent(i) < i


PLUGIN_SRC = r"""
#include "tmkit_b200/amino_rx.h"
/* what an aarxc-generated scene plugin looks like (Makefile.am:213-231; loaded at demo/tmpexec.c:172) */
struct aa_rx_sg *aa_rx_dl_sg__tiny(struct aa_rx_sg *sg) {
    static const double q[4] = {0, 0, 0, 1}, v[3] = {0, 0, 0.5}, ax[3] = {0, 0, 1}, dim[3] = {0.2, 0.2, 0.2};
    aa_rx_sg_add_frame_fixed(sg, "", "base", q, v);
    aa_rx_sg_add_frame_revolute(sg, "base", "j0", q, v, "j0", ax, 0.0);
    aa_rx_sg_set_limit_pos(sg, "j0", -1.0, 2.0);
    aa_rx_geom_attach(sg, "base", aa_rx_geom_box(0, dim));
    aa_rx_geom_attach(sg, "j0", aa_rx_geom_cylinder(0, 0.3, 0.05));
    aa_rx_sg_allow_collision_name(sg, "base", "j0", 1);
    return sg;
}
"""


def test_scene_plugin_loading(built, tmp_path):
    src = tmp_path / "tiny_scene.c"
    src.write_text(PLUGIN_SRC)
    so = tmp_path / "libtiny_scene.so"
    libdir = os.path.dirname(A.LIB_PATH)
    subprocess.check_call(["gcc", "-shared", "-fPIC", f"-I{ROOT}/include", "-o", str(so), str(src),
                           f"-L{libdir}", "-l:libtmkcl.so", f"-Wl,-rpath,{libdir}"])  # as Makefile.am links -lamino
    g = A.SceneGraph.dl(str(so), "tiny")
    assert g is not None
    g.init()
    g.cl_init()
    assert g.frame_names() == ["base", "j0"] and g.config_names() == ["j0"]
    np.testing.assert_allclose(g.limits([0]), [[-1.0, 2.0]])
    assert A.SceneGraph.dl(str(so), "nope") is None            # NULL + message on stderr (tmpexec.c:173-177)
    assert A.SceneGraph.dl(str(tmp_path / "missing.so"), "tiny") is None


def test_misuse_aborts_loudly(built):
    """Amino asserts on misuse; so does this library (message on stderr, abort)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tmkit_b200 import amino as A\n"
            "sg = A.SceneGraph()\n"
            "A.lib().aa_rx_sg_frame_count(sg.h)\n") % ROOT
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "tmkcl:" in r.stderr


def test_no_cpu_fallback_without_a_device(built):
    """compute entry points need a CUDA device and say so instead of computing on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from tmkit_b200 import amino as A, scenes as S\n"
            "sg = A.SceneGraph.from_scene(S.baxter_sussman())\n"
            "sg.tf(np.zeros(sg.config_count()))\n") % ROOT
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
