"""Tuning probe: how much of the deferred narrowphase can hide under the next batch's tile kernel?  Two collision
contexts on two streams take alternate batches (each has its own scratch), against one context taking all."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tmkit_b200 import amino as A, scenes as S, workloads as W

wl = sys.argv[1] if len(sys.argv) > 1 else "configs"
sc = S.baxter_sussman() if wl == "configs" else S.blocksworld_scene(32, seed=32)
sg = A.SceneGraph.from_scene(sc)
q0 = S.q0_vector(sc)
sg.allow_config(q0)
names = S.RIGHT_ARM
sub = sg.config_indices(names)
lim = W.arm_limit_table(names)
n = 1 << 20
RING = 8
dq = [torch.from_numpy(np.ascontiguousarray(W.random_configs(lim, n, seed=r).T.astype(np.float32))).cuda() for r in range(RING)]
bits = [torch.zeros(n // 32, dtype=torch.int32, device="cuda") for _ in range(4)]
for n_ctx in (1, 2):
    cls = [A.Collision(sg) for _ in range(n_ctx)]
    streams = [torch.cuda.Stream(priority=-1 if k else 0) for k in range(n_ctx)]
    def step(i):
        c = i % n_ctx
        cls[c].check_batch_dev(n, sub, q0, dq[i % RING].data_ptr(), A.Q_SOA_F32, n, bits[i % 4].data_ptr(), streams[c].cuda_stream)
    for i in range(4 * RING):
        step(i)
    torch.cuda.synchronize()
    steps = 80
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_event(e0)
    for i in range(steps):
        step(i)
    for s in streams[1:]:
        ev = torch.cuda.Event(); ev.record(s); streams[0].wait_event(ev)
    e1.record(streams[0])
    torch.cuda.synchronize()
    print(f"{wl}: {n_ctx} context(s): {e0.elapsed_time(e1) / steps:.4f} ms per 2^20")
