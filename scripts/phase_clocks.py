"""tuning aid: cycles per pipeline phase (needs a library built with -DTMK_PHASE_CLOCKS, TMK_LIB=...)"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.')
from tmkit_b200 import amino as A, scenes as S, workloads as W
import torch
wl = sys.argv[1] if len(sys.argv) > 1 else "configs"
sc = {"configs": S.baxter_sussman, "blocks": lambda: S.blocksworld_scene(32, seed=32),
      "clutter": lambda: S.clutter_scene(512, seed=0)}[wl]()
names_q = S.RIGHT_ARM + S.LEFT_ARM if wl == "clutter" else S.RIGHT_ARM
sg = A.SceneGraph.from_scene(sc); q0 = S.q0_vector(sc); sg.allow_config(q0); cl = A.Collision(sg)
sub = sg.config_indices(names_q); lim = W.arm_limit_table(names_q)
n = 1 << 20
Q = torch.from_numpy(np.ascontiguousarray(W.random_configs(lim, n, 0).T.astype(np.float32))).cuda()
bits = torch.zeros(n // 32, dtype=torch.int32, device="cuda")
L = A.lib(); L.tmk_phase_cycles.argtypes = [C.c_void_p, C.c_int]
out = (C.c_ulonglong * 8)()
for rep in range(3):
    cl.check_batch_dev(n, sub, q0, Q.data_ptr(), A.Q_SOA_F32, n, bits.data_ptr(), 0)
    cl.sync()
    L.tmk_phase_cycles(out, 1)
v = np.array(list(out)[:6], dtype=float)
names = ["load q", "P1+P2", "Q", "-", "export", "emit"]
print({k: f"{100*x/v.sum():.1f}%" for k, x in zip(names, v)}, "cycles per tile per CTA:", v.sum() / (n / 256))
