"""Worker of tests/test_gpu_multi.py::test_one_process_per_gpu_gathered_bitmap_equals_one_gpu (launched by
torchrun, one rank per GPU).  torch.distributed only carries the 128-byte communicator id to the ranks; the
gather itself is the library's (ncclAllGather inside aa_rx_cl_check_*_shard_dev)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tmkit_b200 import amino as A  # noqa: E402
from tmkit_b200 import scenes as S  # noqa: E402
from tmkit_b200 import workloads as W  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ids = [A.ShardComm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = A.ShardComm(rank, world, ids[0])
    dev = torch.device("cuda", local)

    sc = S.baxter_sussman()
    sg = A.SceneGraph.from_scene(sc)
    q0 = S.q0_vector(sc)
    sg.allow_config(q0)
    cl = A.Collision(sg)
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    seg = W.seg_len(lim)
    s = torch.cuda.Stream()
    res = {"world": world}

    def unpack(t, n):
        return np.unpackbits(t.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)

    n = (1 << 19) + 4321  # ragged: the last slice is short
    Q = W.random_configs(lim, n, seed=31)  # every rank draws the same batch and keeps its slice
    lo, hi = A.shard_bounds(n, world, rank)
    wpr = A.shard_words(n, world)
    sl = torch.from_numpy(np.ascontiguousarray(Q[lo:hi].T.astype(np.float32))).to(dev) if hi > lo else torch.zeros(8, device=dev)
    allb = torch.full((wpr * world,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    ok = True
    for mode in (0, 1):
        comm.overlap(bool(mode))
        for rep in range(3):
            allb.fill_(0x5A5A5A5A)
            torch.cuda.synchronize()
            comm.check_batch_dev(cl, n, sub, q0, sl.data_ptr(), A.Q_SOA_F32, max(1, hi - lo), allb.data_ptr(), s.cuda_stream)
            comm.join(s.cuda_stream)
            s.synchronize()
            if rank == 0:
                want = cl.check_batch(sub, q0, Q)
                ok = ok and bool(np.array_equal(unpack(allb, n), want))
        res["configs_equal" if mode == 0 else "overlap_equal"] = ok
    comm.overlap(False)

    ne = 40_000 + 123
    a, b = W.random_edges(lim, ne, seed=32)
    lo, hi = A.shard_bounds(ne, world, rank)
    wpr = A.shard_words(ne, world)
    da = torch.from_numpy(np.ascontiguousarray(a[lo:hi].T.astype(np.float32))).to(dev)
    db = torch.from_numpy(np.ascontiguousarray(b[lo:hi].T.astype(np.float32))).to(dev)
    eb = torch.zeros(wpr * world, dtype=torch.int32, device=dev)
    ef = torch.zeros(max(1, hi - lo), dtype=torch.int32, device=dev)
    comm.check_edges_dev(cl, ne, sub, q0, da.data_ptr(), db.data_ptr(), A.Q_SOA_F32, hi - lo, seg, eb.data_ptr(), ef.data_ptr(),
                         s.cuda_stream)
    s.synchronize()
    if rank == 0:
        wv, wf = cl.check_edges(sub, q0, a, b, seg)
        res["edges_equal"] = bool(np.array_equal(unpack(eb, ne), wv)) and bool(np.array_equal(ef.cpu().numpy(), wf[lo:hi]))
        with open(sys.argv[1], "w") as f:
            json.dump(res, f)
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
