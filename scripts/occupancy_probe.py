"""Tuning aid: tile-kernel throughput on sub-chains of the right arm (fewer moving geometries -> smaller pose
slab -> more CTAs per SM fit).  Build variants with `make -C tmkit_b200/csrc EXTRA=-DTMK_TILE_MINCTAS\\(T\\)=3 OUT=...`
and point TMK_LIB at them."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from tmkit_b200 import amino as A, scenes as S, workloads as W

sc = S.baxter_sussman()
sg = A.SceneGraph.from_scene(sc)
q0 = S.q0_vector(sc)
sg.allow_config(q0)
for k in (7, 5, 4, 3, 2):
    names = S.RIGHT_ARM[-k:]
    cl = A.Collision(sg)
    sub = sg.config_indices(names)
    lim = W.arm_limit_table(names)
    n = 1 << 20
    Q = W.random_configs(lim, n, seed=k)
    dq = torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).cuda()
    bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device='cuda')
    ts = torch.cuda.Stream()
    with torch.cuda.stream(ts):
        for _ in range(3):
            cl.check_batch_dev(n, sub, q0, dq.data_ptr(), A.Q_SOA_F32, n, bits.data_ptr(), ts.cuda_stream)
        torch.cuda.synchronize()
        cl.kernel_timing(True)
        for _ in range(20):
            cl.check_batch_dev(n, sub, q0, dq.data_ptr(), A.Q_SOA_F32, n, bits.data_ptr(), ts.cuda_stream)
        torch.cuda.synchronize()
        nt, ms, rms = cl.kernel_time()
    print(f"joints {k}: main kernels {ms / nt:.4f} ms, re-check {rms / nt:.4f} ms", flush=True)
    del cl
