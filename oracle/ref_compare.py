"""TEST INFRASTRUCTURE: diff the oracle's verdicts with a REAL Amino / FCL check_state loop (oracle/_ref), when one
can be built on this box (`make -C oracle ref`).  Prints one JSON line; {"amino_fcl": "absent"} otherwise.

    python oracle/ref_compare.py [n_configs]

The scene goes to Amino the way the reference's build does it: as a compiled scene plugin (Makefile.am:213-231),
here generated from the neutral description by tmkit_b200.scenes.to_plugin_source and compiled against Amino's own
headers - so this also checks the restated constructor signatures.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    subprocess.run(["make", "-C", HERE, "ref"], capture_output=True)
    exe = os.path.join(HERE, "_ref", "check_state_ref")
    if not os.path.exists(exe):
        print(json.dumps({"amino_fcl": "absent", "probe": "pkg-config --exists amino amino-collision"}))
        return
    from oracle import oracle as O
    from tmkit_b200 import scenes as S
    from tmkit_b200 import workloads as W
    sc = S.baxter_sussman()
    with tempfile.TemporaryDirectory() as d:
        src, so = os.path.join(d, "scene.c"), os.path.join(d, "scene.so")
        open(src, "w").write(S.to_plugin_source(sc, "sussman"))
        flags = subprocess.run(["pkg-config", "--cflags", "--libs", "amino"], capture_output=True, text=True).stdout.split()
        r = subprocess.run(["gcc", "-O1", "-fPIC", "-shared", src, "-o", so, *flags], capture_output=True, text=True)
        if r.returncode:
            print(json.dumps({"amino_fcl": "present", "plugin_build": r.stderr[-600:]}))
            return
        orc = O.OracleScene(S.flatten(sc))
        q0 = S.q0_vector(sc)
        lim = W.arm_limit_table(S.RIGHT_ARM)
        sub = orc.config_ids(S.RIGHT_ARM)
        Q = np.repeat(q0[None, :], n + 1, axis=0)
        Q[1:, sub] = W.random_configs(lim, n, seed=0)
        cfg, out = os.path.join(d, "configs.bin"), os.path.join(d, "verdicts.bin")
        with open(cfg, "wb") as f:
            np.asarray([n + 1, Q.shape[1]], dtype=np.uint64).tofile(f)
            Q.tofile(f)
        r = subprocess.run([exe, so, "sussman", cfg, out], capture_output=True, text=True)
        if r.returncode:
            print(json.dumps({"amino_fcl": "present", "run": r.stderr[-600:]}))
            return
        ref = np.fromfile(out, dtype=np.uint8)[1:].astype(bool)
        orc.allow_config(q0)
        st, _ = orc.check_batch(sub, q0, Q[1:, sub], threads=os.cpu_count() or 1)
        decided = st != O.BAND
        mism = int((ref[decided] != (st[decided] == O.HIT)).sum())
        print(json.dumps({"amino_fcl": "present", "configs": n, "band": int((~decided).sum()), "mismatches": mism,
                          "seconds_per_check_amino": float(r.stdout.strip() or "nan")}))


if __name__ == "__main__":
    main()
