"""tuning aid: host spans of the blocking host-buffer call (TMK_PROFILE_HOST=1), double and float rows"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from tmkit_b200 import amino as A, scenes as S, workloads as W
sc = S.baxter_sussman(); sg = A.SceneGraph.from_scene(sc); q0 = S.q0_vector(sc); sg.allow_config(q0); cl = A.Collision(sg)
sub = sg.config_indices(S.RIGHT_ARM); lim = W.arm_limit_table(S.RIGHT_ARM)
n = 1 << 20
Q = W.random_configs(lim, n, seed=0)
h64 = torch.from_numpy(Q).pin_memory(); h32 = torch.from_numpy(Q.astype(np.float32)).pin_memory()
hb = np.zeros(n // 32 + 1, dtype=np.uint32)
for name, fn, h in (("f64", cl.L.aa_rx_cl_check_batch, h64), ("f32", cl.L.aa_rx_cl_check_batch_f32, h32)):
    for rep in range(6):
        t = time.perf_counter()
        fn(cl.h, n, len(sub), sub.ctypes.data, len(q0), q0.ctypes.data, h.data_ptr(), len(sub), hb.ctypes.data)
        print(name, rep, f"{(time.perf_counter() - t) * 1e3:.3f} ms", file=sys.stderr, flush=True)
