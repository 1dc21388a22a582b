"""Generates tests/golden/sussman_golden.npz from the CPU oracle (run once, commit the output).

The reference tree holds no golden vector for this path (no TESTS in
Makefile.am), so the frozen vectors come from the oracle itself; what pins the oracle to the
reference is (a) the world poses derived by hand from the reference's scene files
(asserted in tests/test_oracle_golden.py) and (b) the reference's own
runtime self-checks (demo/tmpexec.c:424: after allow_config(q), check(q) == 0).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tmkit_b200 import scenes as S  # noqa: E402
from tmkit_b200 import workloads as W  # noqa: E402


def main():
    sc = S.baxter_sussman()
    flat = S.flatten(sc)
    orc = O.OracleScene(flat)
    q0 = S.q0_vector(sc)
    rel0, abs0 = orc.fk(q0)
    st0, ex0, ps0 = orc.check(abs0)
    orc.allow_config(q0)
    sub = orc.config_ids(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    out = {"names": np.array(orc.names), "config_names": np.array(orc.config_names), "q0": q0,
           "tf_rel_q0": rel0, "tf_abs_q0": abs0, "pairset_q0": ps0, "allowed_after_q0": orc.allowed.copy()}
    for seed in (0, 1, 2):
        Q = W.random_configs(lim, 4096, seed)
        st, ex = orc.check_batch(sub, q0, Q, threads=os.cpu_count() or 1)
        out[f"cfg_state_{seed}"] = st
        out[f"cfg_exact_{seed}"] = ex
        a, b = W.random_edges(lim, 512, seed)
        est, eex, first, nchk = orc.check_edges(sub, q0, a, b, W.seg_len(lim), threads=os.cpu_count() or 1)
        out[f"edge_state_{seed}"] = est
        out[f"edge_first_{seed}"] = first
    # FK golden at a few random full configurations
    rng = np.random.default_rng(42)
    qs = rng.uniform(-1.5, 1.5, size=(8, orc.n_q))
    out["fk_q"] = qs
    out["fk_abs"] = np.stack([orc.fk(q)[1] for q in qs])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sussman_golden.npz"), **out)
    print("wrote golden:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
