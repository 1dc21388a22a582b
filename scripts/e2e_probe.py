"""Tuning probe: the blocking host-buffer call (aa_rx_cl_check_batch) on 2^20 pinned double rows; prints ms per call
for float / double rows.  TMK_NO_GROUPS=1 / TMK_GROUP_WAVES=n select the pipeline variant; TMK_PROFILE_HOST=1 prints spans."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tmkit_b200 import amino as A, scenes as S, workloads as W

wl = sys.argv[1] if len(sys.argv) > 1 else "configs"
sc = S.baxter_sussman() if wl == "configs" else S.blocksworld_scene(32, seed=32)
sg = A.SceneGraph.from_scene(sc)
q0 = S.q0_vector(sc)
sg.allow_config(q0)
cl = A.Collision(sg)
sub = sg.config_indices(S.RIGHT_ARM)
lim = W.arm_limit_table(S.RIGHT_ARM)
n = 1 << 20
Q = [torch.from_numpy(W.random_configs(lim, n, seed=r)).pin_memory() for r in range(2)]
Q32 = [torch.from_numpy(q.numpy().astype(np.float32)).pin_memory() for q in Q]
hb = np.zeros(n // 32 + 1, dtype=np.uint32)
ref = None
for name, rows, fn in (("f64", Q, cl.L.aa_rx_cl_check_batch), ("f32", Q32, cl.L.aa_rx_cl_check_batch_f32)):
    for i in range(3):
        fn(cl.h, n, len(sub), sub.ctypes.data, len(q0), q0.ctypes.data, rows[i % 2].data_ptr(), len(sub), hb.ctypes.data)
    t0 = time.perf_counter()
    for i in range(20):
        fn(cl.h, n, len(sub), sub.ctypes.data, len(q0), q0.ctypes.data, rows[i % 2].data_ptr(), len(sub), hb.ctypes.data)
    dt = (time.perf_counter() - t0) / 20
    print(f"{wl} {name} rows: {dt * 1e3:.3f} ms per 2^20  ({n / dt:.3e} configs/s)  checksum {int(hb[:n // 32].astype(np.uint64).sum())}")
