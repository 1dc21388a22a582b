"""Multi-GPU entry points of the C ABI (include/tmkit_b200/amino_rx.h, "Multi-GPU"): slices, the gathered
bitmap, and the exact code path bench.py times.

The contract is the same as for one GPU - verdicts are a pure function of the inputs - so every multi-device
result must EQUAL the single-device bitmap bit for bit ("comparing the gathered bitmap against the
single-device result").  Tests that need two devices skip on a one-GPU box; the slicing logic itself is also run
on one GPU (two contexts on the same device).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tmkit_b200 import amino as A
from tmkit_b200 import scenes as S
from tmkit_b200 import workloads as W

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORES = os.cpu_count() or 1


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def arm(sussman):
    sc, flat, orc, q0 = sussman
    sg = A.SceneGraph.from_scene(sc)
    sg.allow_config(q0)
    cl = A.Collision(sg)
    sub = sg.config_indices(S.RIGHT_ARM)
    lim = W.arm_limit_table(S.RIGHT_ARM)
    return sc, orc, sg, cl, q0, sub, lim


def _unpack(words, n):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n].astype(bool)


# ------------------------------------------------------------------ no write past ceil(n/32) words
@pytest.mark.parametrize("n", [1, 31, 100, 257, 1000, 5000])
def test_dev_bitmap_is_not_written_past_its_last_word(arm, n):
    import torch
    sc, orc, sg, cl, q0, sub, lim = arm
    Q = W.random_configs(lim, n, seed=n)
    want = cl.check_batch(sub, q0, Q)
    nw = (n + 31) // 32
    dq = torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).cuda()
    canary = 0x5A5A5A5A
    bits = torch.full((nw + 64,), canary, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):  # third call replays the library's CUDA graph
        cl.check_batch_dev(n, sub, q0, dq.data_ptr(), A.Q_SOA_F32, n, bits.data_ptr(), s)
    torch.cuda.synchronize()
    h = bits.cpu().numpy().view(np.uint32)
    assert (h[nw:] == canary).all(), "words past ceil(n/32) were overwritten"
    assert np.array_equal(_unpack(h[:nw], n), want)
    a, b = W.random_edges(lim, n, seed=n)
    ev, _ = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    da = torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).cuda()
    db = torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).cuda()
    bits.fill_(canary)
    for _ in range(2):
        cl.check_edges_dev(n, sub, q0, da.data_ptr(), db.data_ptr(), A.Q_SOA_F32, n, W.seg_len(lim), bits.data_ptr(), 0, s)
    torch.cuda.synchronize()
    h = bits.cpu().numpy().view(np.uint32)
    assert (h[nw:] == canary).all()
    assert np.array_equal(_unpack(h[:nw], n), ev)


# ------------------------------------------------------------------ edges deeper than the bisection levels
def test_edges_with_more_segments_than_the_level_pipeline_covers(arm):
    """seg_len so small that nd exceeds 2^13: the level pipeline cannot reach every position; the edge must still be
    decided exactly as OMPL's order does (the overflow path walks all states)."""
    sc, orc, sg, cl, q0, sub, lim = arm
    a, b = W.random_edges(lim, 24, seed=5)
    seg = 2.0e-4  # the reference's 1 % resolution is 0.124; edges here are up to 2.5 rad long -> nd up to ~12 000
    nd = np.ceil(np.linalg.norm(b - a, axis=1) / seg)
    assert nd.max() > (1 << 13)
    valid, first = cl.check_edges(sub, q0, a, b, seg)
    est, _, efirst, _ = orc.check_edges(orc.config_ids(S.RIGHT_ARM), q0, a, b, seg, threads=CORES)
    dec = est != O.BAND
    assert np.array_equal(~valid[dec], est[dec] == O.HIT)
    hit = dec & (est == O.HIT)
    assert np.array_equal(first[hit], efirst[hit])
    # the same through the device-pointer entry (all levels launched, no host-side level clamp)
    import torch
    da = torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).cuda()
    db = torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).cuda()
    bits = torch.zeros(1, dtype=torch.int32, device="cuda")
    fi = torch.zeros(24, dtype=torch.int32, device="cuda")
    cl.check_edges_dev(24, sub, q0, da.data_ptr(), db.data_ptr(), A.Q_SOA_F32, 24, seg, bits.data_ptr(), fi.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(_unpack(bits.cpu().numpy().view(np.uint32), 24), valid)
    assert np.array_equal(fi.cpu().numpy(), first)


# ------------------------------------------------------------------ one process, several contexts
@pytest.mark.parametrize("n", [0, 1, 1000, 1024, 5000, 70_001])
def test_multi_host_calls_on_one_device_equal_the_single_call(arm, n):
    """two and three slices on device 0 (two contexts on one GPU): slice offsets, word offsets, short last slice"""
    sc, orc, sg, cl, q0, sub, lim = arm
    Q = W.random_configs(lim, n, seed=11)
    want = cl.check_batch(sub, q0, Q)
    a, b = W.random_edges(lim, min(n, 6000), seed=12)
    ev, ef = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    for devices in ([0, 0], [0, 0, 0]):
        m = A.CollisionMulti(sg, devices)
        assert m.n_dev == len(devices)
        assert np.array_equal(m.check_batch(sub, q0, Q), want)
        assert np.array_equal(m.check_batch(sub, q0, Q.astype(np.float32)), want)
        gv, gf = m.check_edges(sub, q0, a, b, W.seg_len(lim))
        assert np.array_equal(gv, ev) and np.array_equal(gf, ef)


def test_multi_allow_and_collision_feedback(arm):
    sc, orc, sg, cl, q0, sub, lim = arm
    Q = W.random_configs(lim, 4000, seed=13)
    m = A.CollisionMulti(sg, [0, 0])
    cl2 = A.Collision(sg)
    cl2.track_collisions(True)
    cl2.check_batch(sub, q0, Q)
    want = cl2.batch_collisions().matrix()
    m.track_collisions(True)
    m.check_batch(sub, q0, Q)
    assert np.array_equal(m.batch_collisions().matrix(), want)
    # allow every pair that showed up: the whole batch becomes valid on every slice
    m.track_collisions(False)
    m.allow_set(cl2.batch_collisions())
    assert m.check_batch(sub, q0, Q).all()


@pytest.mark.skipif("_n_gpus() < 2")
@pytest.mark.parametrize("n", [5000, 1 << 20, (1 << 20) + 77])
def test_two_gpus_equal_one_gpu(arm, n):
    """BASELINE configs[4] contract: the gathered multi-GPU bitmap equals the 1-GPU bitmap bit for bit"""
    import torch
    sc, orc, sg, cl, q0, sub, lim = arm
    world = min(_n_gpus(), 8)
    Q = W.random_configs(lim, n, seed=21)
    want = cl.check_batch(sub, q0, Q)
    m = A.CollisionMulti(sg, list(range(world)))
    assert np.array_equal(m.check_batch(sub, q0, Q), want)       # host rows, pageable
    a, b = W.random_edges(lim, min(n, 50_000), seed=22)
    ev, ef = cl.check_edges(sub, q0, a, b, W.seg_len(lim))
    gv, gf = m.check_edges(sub, q0, a, b, W.seg_len(lim))
    assert np.array_equal(gv, ev) and np.array_equal(gf, ef)
    # device-resident: every device holds its slice (float rows) and a copy of the gathered bitmap; the words
    # are completed by the library's ncclAllGather
    wpr = A.shard_words(n, world)
    soa = n % (world * A.SHARD_ALIGN) == 0  # equal slices: the joint-major layout has one ld for all of them
    dQ, dB = [], []
    for k in range(world):
        lo, hi = A.shard_bounds(n, world, k)
        dev = torch.device("cuda", k)
        rows = Q[lo:hi].astype(np.float32) if hi > lo else np.zeros((1, len(sub)), np.float32)
        dQ.append(torch.from_numpy(np.ascontiguousarray(rows.T if soa else rows)).to(dev))
        dB.append(torch.full((wpr * world,), 0x5A5A5A5A, dtype=torch.int32, device=dev))
    for _ in range(3):  # the third call replays the per-device CUDA graphs
        if soa:
            m.check_batch_dev(n, sub, q0, [t.data_ptr() for t in dQ], A.Q_SOA_F32, n // world, [t.data_ptr() for t in dB])
        else:
            m.check_batch_dev(n, sub, q0, [t.data_ptr() for t in dQ], A.Q_ROWS_F32, len(sub), [t.data_ptr() for t in dB])
        m.sync()
        for k in range(world):
            got = _unpack(dB[k].cpu().numpy().view(np.uint32), n)
            assert np.array_equal(got, want), f"device {k}"


@pytest.mark.skipif("_n_gpus() < 2")
def test_one_process_per_gpu_gathered_bitmap_equals_one_gpu(tmp_path):
    """torchrun, one rank per GPU, tmk_shard_comm + aa_rx_cl_check_batch_shard_dev: rank 0 compares the gathered
    bitmap with the bitmap it computes alone"""
    world = min(_n_gpus(), 8)
    out = tmp_path / "shard.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29655", os.path.join(ROOT, "tests", "shard_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    res = json.loads(out.read_text())
    assert res["world"] == world and res["configs_equal"] and res["edges_equal"] and res["overlap_equal"], res


# ------------------------------------------------------------------ the loop bench.py times, checked
def test_bench_loop_configs_replayed_graphs_match_host_call_and_oracle(arm):
    """aa_rx_cl_check_batch_dev, SoA float32, 2^20 per batch, a ring of inputs, enough passes that the library
    replays its CUDA graphs: every pass equals the host-buffer call bit for bit and the oracle on a sample"""
    import torch
    sc, orc, sg, cl, q0, sub, lim = arm
    n, ring = 1 << 20, 3
    Qs = [W.random_configs(lim, n, seed=100 + r) for r in range(ring)]
    want = [cl.check_batch(sub, q0, Q) for Q in Qs]
    dq = [torch.from_numpy(np.ascontiguousarray(Q.T.astype(np.float32))).cuda() for Q in Qs]
    nw = n // 32
    bits = [torch.zeros(nw, dtype=torch.int32, device="cuda") for _ in range(2)]
    ts = torch.cuda.Stream()
    osub = orc.config_ids(S.RIGHT_ARM)
    launches = []
    with torch.cuda.stream(ts):
        for i in range(4 * ring):
            k, b = i % ring, i & 1
            bits[b].fill_(0x5A5A5A5A)
            cl.check_batch_dev(n, sub, q0, dq[k].data_ptr(), A.Q_SOA_F32, n, bits[b].data_ptr(), ts.cuda_stream)
            launches.append(int(cl.stats().n_launches))
            ts.synchronize()
            got = _unpack(bits[b].cpu().numpy().view(np.uint32), n)
            assert np.array_equal(got, want[k]), f"pass {i} (ring {k})"
    assert launches[-1] > 0
    for k in range(ring):
        idx = np.random.default_rng(k).choice(n, 20000 // ring, replace=False)
        st, _ = orc.check_batch(osub, q0, Qs[k][idx], threads=CORES, early_out=True)
        dec = st != O.BAND
        assert np.array_equal(~want[k][idx][dec], st[dec] == O.HIT)


def test_bench_loop_edges_replayed_graphs_match_host_call_and_oracle(arm):
    import torch
    sc, orc, sg, cl, q0, sub, lim = arm
    n, ring = 100_000, 2
    seg = W.seg_len(lim)
    E = [W.random_edges(lim, n, seed=200 + r) for r in range(ring)]
    want = [cl.check_edges(sub, q0, a, b, seg) for a, b in E]
    da = [torch.from_numpy(np.ascontiguousarray(a.T.astype(np.float32))).cuda() for a, _ in E]
    db = [torch.from_numpy(np.ascontiguousarray(b.T.astype(np.float32))).cuda() for _, b in E]
    nw = (n + 31) // 32
    bits = torch.zeros(nw, dtype=torch.int32, device="cuda")
    first = torch.zeros(n, dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    with torch.cuda.stream(ts):
        for i in range(4 * ring):
            k = i % ring
            bits.fill_(0x5A5A5A5A)
            cl.check_edges_dev(n, sub, q0, da[k].data_ptr(), db[k].data_ptr(), A.Q_SOA_F32, n, seg, bits.data_ptr(),
                               first.data_ptr(), ts.cuda_stream)
            ts.synchronize()
            assert np.array_equal(_unpack(bits.cpu().numpy().view(np.uint32), n), want[k][0]), f"pass {i}"
            assert np.array_equal(first.cpu().numpy(), want[k][1]), f"pass {i}"
    osub = orc.config_ids(S.RIGHT_ARM)
    idx = np.random.default_rng(9).choice(n, 3000, replace=False)
    a, b = E[0]
    est, _, efirst, _ = orc.check_edges(osub, q0, a[idx], b[idx], seg, threads=CORES)
    dec = est != O.BAND
    assert np.array_equal(~want[0][0][idx][dec], est[dec] == O.HIT)
