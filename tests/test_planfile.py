"""Plan files (N2): writer / parser of the reference's line format (doc/md/planfile.md:13-76,
src/planfile.l:77-192).  No GPU needed."""
import numpy as np
import pytest

from tmkit_b200 import amino as A
from tmkit_b200 import scenes as S

# demo/baxter-sussman/q0.tmp:1-2 verbatim (data; the start configuration file of the named benchmark)
Q0_TMP = ("m right_w2 right_w1 right_w0 right_s1 right_s0 right_e1 right_e0 left_w2 left_w1 left_w0 left_s1 left_s0 "
          "left_e1 left_e0\n"
          "p 0.0 1.5707963267948966 0.0 -0.7853981633974483 0.1570796350201586 0.7853981633974483 0.0 0.0 "
          "0.7853981633974483 0.0 -1.5707963267948966 2.04203514993196 2.356194490192345 0.0\n")


def test_parse_reference_start_file(built, tmp_path):
    f = tmp_path / "q0.tmp"
    f.write_text(Q0_TMP)
    plan = A.PlanFile.parse(str(f))
    ops = plan.ops()
    assert len(ops) == 1 and ops[0][0] == "m"
    names, data = ops[0][1], ops[0][2]
    assert data.shape == (1, 14)
    # map by NAME (demo/tmpexec.c:269-272): equals the fixture's start configuration
    assert dict(zip(names, data[0])) == S.Q0


def test_round_trip_is_exact(built, tmp_path):
    rng = np.random.default_rng(0)
    plan = A.PlanFile()
    plan.add_action("  UNSTACK block_c block_a ")
    pts = rng.normal(size=(5, 7)) * np.pi
    plan.add_motion(S.RIGHT_ARM, pts)
    plan.add_reparent("block_c", "end_effector_grasp")
    plan.add_action("PUT-DOWN block_c front_table 2 -1")
    wide = np.full((3, 9), np.nan)
    wide[:, :7] = rng.normal(size=(3, 7))
    plan.add_motion(S.RIGHT_ARM, wide)  # leading dimension > number of names
    plan.add_reparent("block_c", "front_table")
    f = tmp_path / "plan.tmp"
    assert plan.write(str(f)) == 0
    text = f.read_text().splitlines()
    assert text[0] == "a UNSTACK block_c block_a"
    assert text[1] == "m " + " ".join(S.RIGHT_ARM)
    assert text[2].startswith("p ") and len(text[2].split()) == 8
    assert text[7] == "r block_c end_effector_grasp"
    back = A.PlanFile.parse(str(f)).ops()
    assert [o[0] for o in back] == ["a", "m", "r", "a", "m", "r"]
    assert back[0][1] == "UNSTACK block_c block_a"
    assert np.array_equal(back[1][2], pts)  # %.17g round-trips doubles exactly
    assert np.array_equal(back[4][2], wide[:, :7])
    assert back[5] == ("r", "block_c", "front_table")


def test_comments_blank_lines_and_errors(built, tmp_path):
    f = tmp_path / "c.tmp"
    f.write_text("# a comment\n\n a PICK-UP block_a   # trailing comment\nm j0 j1\np 1 2\n  p 3 4 # second\n")
    ops = A.PlanFile.parse(str(f)).ops()
    assert ops[0] == ("a", "PICK-UP block_a")
    assert np.array_equal(ops[1][2], [[1, 2], [3, 4]])
    for bad in ("p 1 2\n", "m j0 j1\np 1\n", "m j0\np 1 2\n", "x what\n", "m\n", "r only_one\n", "m j0\np abc\n"):
        g = tmp_path / "bad.tmp"
        g.write_text(bad)
        assert A.PlanFile.parse(str(g)) is None, bad
    assert A.PlanFile.parse(str(tmp_path / "missing.tmp")) is None


# ------------------------------------------------------------------ the reference's own names (libtmsmt's tmplan_*)
import ctypes as C


def _tmplan():
    """bind the tmplan_* symbols of include/tmkit_b200/tmplan.h (= /root/reference/include/tmsmt/tmplan.h:34-247)"""
    L = A.lib()
    vp, sz, cs = C.c_void_p, C.c_size_t, C.c_char_p
    sig = {
        "tmplan_create": (vp, [vp]), "tmplan_destroy": (None, [vp]), "tmplan_op_count": (sz, [vp]),
        "tmplan_finish_op": (None, [vp]), "tmplan_add_action": (None, [vp]), "tmplan_add_motion_plan": (None, [vp]),
        "tmplan_add_reparent": (None, [vp]), "tmplan_last_op": (vp, [vp]), "tmplan_op_type": (C.c_int, [vp]),
        "tmplan_op_action_set": (None, [vp, cs]), "tmplan_op_action_get": (cs, [vp]),
        "tmplan_op_motion_plan_add_var": (None, [vp, cs]), "tmplan_op_motion_plan_map_var": (None, [vp, vp, vp]),
        "tmplan_op_motion_plan_vars": (None, [vp, sz, vp]), "tmplan_map_ops": (C.c_int, [vp, vp, vp]),
        "tmplan_op_motion_plan_path_start": (None, [vp]), "tmplan_op_motion_plan_path_add": (None, [vp, C.c_double]),
        "tmplan_op_motion_plan_path_finish": (None, [vp]), "tmplan_op_motion_plan_path": (vp, [vp]),
        "tmplan_op_motion_plan_config_count": (sz, [vp]), "tmplan_op_motion_plan_path_size": (sz, [vp]),
        "tmplan_op_reparent_set_frame": (None, [vp, cs]), "tmplan_op_reparent_set_new_parent": (None, [vp, cs]),
        "tmplan_op_reparent_get_frame": (cs, [vp]), "tmplan_op_reparent_get_new_parent": (cs, [vp]),
        "tmplan_parse_filename": (vp, [cs, vp]), "tmplan_parse_file": (vp, [vp, vp]), "tmplan_write": (C.c_int, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


OP_ACTION, OP_MOTION_PLAN, OP_REPARENT = 0, 1, 2  # enum tmplan_op_type (tmplan.h:13-17)
OP_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
VAR_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)


def _walk(L, plan):
    """what lisp/foreign-tmplan.lisp:123-153 does: tmplan_map_ops with a callback that switches on the op type"""
    out = []

    def on_op(_cx, op):
        t = L.tmplan_op_type(op)
        if t == OP_ACTION:
            out.append(("a", L.tmplan_op_action_get(op).decode()))
        elif t == OP_REPARENT:
            out.append(("r", L.tmplan_op_reparent_get_frame(op).decode(), L.tmplan_op_reparent_get_new_parent(op).decode()))
        else:
            names = []
            L.tmplan_op_motion_plan_map_var(op, C.cast(VAR_CB(lambda _c, n: names.append(n.decode())), C.c_void_p), None)
            nq, ne = L.tmplan_op_motion_plan_config_count(op), L.tmplan_op_motion_plan_path_size(op)
            arr = (C.c_char_p * nq)()
            L.tmplan_op_motion_plan_vars(op, nq, arr)
            assert [a.decode() for a in arr] == names
            data = np.ctypeslib.as_array(C.cast(L.tmplan_op_motion_plan_path(op), C.POINTER(C.c_double)), shape=(ne,)).copy() \
                if ne else np.zeros(0)
            out.append(("m", names, data.reshape(-1, nq)))

    cb = OP_CB(on_op)
    L.tmplan_map_ops(plan, C.cast(cb, C.c_void_p), None)
    return out


def test_tmplan_abi_parses_the_reference_start_file(built, tmp_path):
    """tmplan_parse_filename + tmplan_map_ops + the accessors demo/tmpexec.c:111-125, 246-312 uses"""
    L = _tmplan()
    f = tmp_path / "q0.tmp"
    f.write_text(Q0_TMP)
    plan = L.tmplan_parse_filename(str(f).encode(), None)
    assert plan and L.tmplan_op_count(plan) == 1
    ops = _walk(L, plan)
    assert ops[0][0] == "m" and dict(zip(ops[0][1], ops[0][2][0])) == S.Q0
    L.tmplan_destroy(plan)
    assert not L.tmplan_parse_filename(str(tmp_path / "missing.tmp").encode(), None)


def test_tmplan_builder_writer_and_the_tmk_plan_api_share_one_format(built, tmp_path):
    """build a plan op by op as the reference's scanner does (src/planfile.l:95-192: add_*, last_op, set / add_var /
    path_add), write it with tmplan_write(FILE*), and read it back with both parsers"""
    L = _tmplan()
    libc = C.CDLL(None)
    libc.fopen.restype, libc.fopen.argtypes = C.c_void_p, [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    plan = L.tmplan_create(None)
    L.tmplan_add_action(plan)
    L.tmplan_op_action_set(L.tmplan_last_op(plan), b"PICK-UP block_a front_table")
    L.tmplan_finish_op(plan)
    L.tmplan_add_motion_plan(plan)
    mp = L.tmplan_last_op(plan)
    for n in ("right_s0", "right_s1", "right_e0"):
        L.tmplan_op_motion_plan_add_var(mp, n.encode())
    L.tmplan_op_motion_plan_path_start(mp)
    pts = np.random.default_rng(3).normal(size=(4, 3))
    for x in pts.ravel():
        L.tmplan_op_motion_plan_path_add(mp, float(x))
    L.tmplan_op_motion_plan_path_finish(mp)
    L.tmplan_add_reparent(plan)
    rp = L.tmplan_last_op(plan)
    L.tmplan_op_reparent_set_frame(rp, b"block_a")
    L.tmplan_op_reparent_set_new_parent(rp, b"end_effector_grasp")
    assert L.tmplan_op_count(plan) == 3
    assert [L.tmplan_op_type(o) for o in (L.tmplan_last_op(plan),)] == [OP_REPARENT]
    path = str(tmp_path / "built.tmp").encode()
    fh = libc.fopen(path, b"w")
    assert L.tmplan_write(plan, fh) == 0
    libc.fclose(fh)
    L.tmplan_destroy(plan)
    text = open(path).read().splitlines()
    assert text[0] == "a PICK-UP block_a front_table" and text[1] == "m right_s0 right_s1 right_e0" and text[-1] == "r block_a end_effector_grasp"
    # read back through the reference-named ABI ...
    back = L.tmplan_parse_filename(path, None)
    ops = _walk(L, back)
    L.tmplan_destroy(back)
    assert ops[0] == ("a", "PICK-UP block_a front_table") and ops[2] == ("r", "block_a", "end_effector_grasp")
    assert ops[1][1] == ["right_s0", "right_s1", "right_e0"] and np.array_equal(ops[1][2], pts)
    # ... and through tmk_plan_parse: one structure behind both
    ops2 = A.PlanFile.parse(path.decode()).ops()
    assert ops2[0] == ops[0] and ops2[2] == ops[2] and np.array_equal(ops2[1][2], pts)
