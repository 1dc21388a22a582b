"""Multi-rank host logic on CPU: world_size-2 (and 4) gloo runs of tmkit_b200.shard.  The
per-slice validity check is played by the CPU oracle here (tests may use it as the checker's
stand-in; on the GPU box the slice is checked by aa_rx_cl_check_batch_dev, see bench.py), so
what is under test is the slicing, the bitmap word layout and the all-gather: the gathered
bitmap must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest

from tmkit_b200 import shard


def test_slice_bounds_properties():
    for n in (0, 1, 31, 1024, 1025, 100_000, 1 << 20, (1 << 24) + 7):
        for world in (1, 2, 3, 4, 8):
            b = shard.slice_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            for (lo, hi), (lo2, hi2) in zip(b, b[1:]):
                assert hi == lo2 and lo <= hi
            for lo, hi in b:
                assert lo % shard.ALIGN == 0 or lo == n
            assert shard.words_per_rank(n, world) * 32 * world >= n
    with pytest.raises(ValueError):
        shard.slice_bounds(10, 0)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    v = rng.random(1000) < 0.5
    w = shard.pack(v, 32)
    assert np.array_equal(shard.unpack(w, 1000), v)
    assert (w[0] >> 3) & 1 == int(v[3])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, kind, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from tmkit_b200 import scenes as S, workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = S.baxter_sussman()
    orc = O.OracleScene(S.flatten(sc))
    q0 = S.q0_vector(sc)
    orc.allow_config(q0)
    sub = orc.config_ids(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    if kind == "configs":
        Q = W.random_configs(lim, n, seed=3)
    else:
        a, b = W.random_edges(lim, n, seed=3)

    def check_slice(lo, hi, words):
        if kind == "configs":
            st, ex = orc.check_batch(sub, q0, Q[lo:hi], early_out=True)
        else:
            st, ex, _, _ = orc.check_edges(sub, q0, a[lo:hi], b[lo:hi], W.seg_len(lim), early_out=True)
        words[:] = torch.from_numpy(shard.pack(~ex.astype(bool), len(words)).view(np.int32))

    words = shard.check_sharded(n, rank, world, check_slice, lambda k: torch.zeros(k, dtype=torch.int32),
                                dist.all_gather_into_tensor)
    np.save(os.path.join(out_dir, f"bits_{rank}.npy"), words.numpy().view(np.uint32))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,kind", [(2, 3000, "configs"), (4, 2500, "configs"), (2, 1100, "edges")])
def test_gather_equals_single_process(tmp_path, world, n, kind):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from tmkit_b200 import scenes as S, workloads as W
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, kind, str(tmp_path)), nprocs=world, join=True)
    sc = S.baxter_sussman()
    orc = O.OracleScene(S.flatten(sc))
    q0 = S.q0_vector(sc)
    orc.allow_config(q0)
    sub = orc.config_ids(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    if kind == "configs":
        st, ex = orc.check_batch(sub, q0, W.random_configs(lim, n, seed=3), early_out=True, threads=4)
    else:
        a, b = W.random_edges(lim, n, seed=3)
        st, ex, _, _ = orc.check_edges(sub, q0, a, b, W.seg_len(lim), early_out=True, threads=4)
    want = ~ex.astype(bool)
    for r in range(world):
        got = shard.unpack(np.load(tmp_path / f"bits_{r}.npy"), n)
        assert np.array_equal(got, want), r
    # every rank holds the same, complete bitmap
    assert len(np.load(tmp_path / "bits_0.npy")) == (n + 31) // 32
