"""Profiling target: a few device-resident steps of one workload, nothing else (short enough for ncu --set full).
usage: prof_step.py [configs|blocks|clutter|edges] [steps] [log2 of the batch size]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tmkit_b200 import amino as A, scenes as S, workloads as W

wl = sys.argv[1] if len(sys.argv) > 1 else "configs"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = {"configs": S.baxter_sussman, "edges": S.baxter_sussman, "blocks": lambda: S.blocksworld_scene(32, seed=32),
      "clutter": lambda: S.clutter_scene(512, seed=0)}[wl]()
names = S.RIGHT_ARM + S.LEFT_ARM if wl == "clutter" else S.RIGHT_ARM
sg = A.SceneGraph.from_scene(sc)
q0 = S.q0_vector(sc)
sg.allow_config(q0)
cl = A.Collision(sg)
sub = sg.config_indices(names)
lim = W.arm_limit_table(names)
n = 100_000 if wl == "edges" else 1 << 20
if len(sys.argv) > 3:
    n = 1 << int(sys.argv[3])
s = torch.cuda.Stream()
bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
if wl == "edges":
    a, b = W.random_edges(lim, n, seed=0)
    da = torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).cuda()
    db = torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).cuda()
    for i in range(steps):
        cl.check_edges_dev(n, sub, q0, da.data_ptr(), db.data_ptr(), A.Q_SOA_F32, n, W.seg_len(lim), bits.data_ptr(), 0, s.cuda_stream)
else:
    dq = torch.from_numpy(np.ascontiguousarray(W.random_configs(lim, n, seed=0).T.astype(np.float32))).cuda()
    for i in range(steps):
        cl.check_batch_dev(n, sub, q0, dq.data_ptr(), A.Q_SOA_F32, n, bits.data_ptr(), s.cuda_stream)
torch.cuda.synchronize()
print("valid", int(np.unpackbits(bits.cpu().numpy().view(np.uint8)).sum()), "of", n)
