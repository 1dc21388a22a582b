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
