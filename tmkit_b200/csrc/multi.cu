/* multi.cu - the validity check over the GPUs of one box (BASELINE.json configs[4]).
 *
 * The path shards: configurations and edges are independent.  Every GPU holds a replica of the scene
 * image (KBs, rebuilt per motion query) and checks one contiguous slice of the batch; slices are aligned
 * to TMK_SHARD_ALIGN units so that every GPU owns whole 32-bit words of the validity bitmap and whole
 * tiles of the pipeline.  The only exchange is the bitmap itself:
 *
 *   host-buffer calls      (aa_rx_cl_check_batch_multi, _edges_multi): one process drives all devices -
 *                          what the reference's single-threaded planner would call (lisp/python/
 *                          tmsmtpy.lisp:55-60; one handle per planner, demo/tmpexec.c:397-399).  Each device
 *                          copies its slice of the caller's rows in, and its words straight back into the
 *                          caller's bitmap: no collective is needed.
 *   device-resident calls  (aa_rx_cl_check_batch_multi_dev: one process, ncclCommInitAll;
 *                          aa_rx_cl_check_batch_shard_dev / _edges_shard_dev: one process per GPU, the
 *                          communicator built from an id the caller distributes): every GPU writes its words
 *                          into its own slot of its copy of the global bitmap and ONE in-place ncclAllGather
 *                          (NVLink / NVSwitch) completes all copies.  NCCL is used for nothing else.
 *
 * No reference counterpart exists: the reference's planner is single-threaded and single-device.
 */
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "tmk_internal.h"

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            tmk_die("CUDA error %s at %s:%d: %s (no CPU fallback exists)", cudaGetErrorName(e_), __FILE__, \
                    __LINE__, cudaGetErrorString(e_));                                               \
    } while (0)
#define NK(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess) tmk_die("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(r_)); \
    } while (0)

/* ------------------------------------------------------------------ slice arithmetic (no device needed) */
extern "C" size_t tmk_shard_per_rank(size_t n, int world) {
    if (world < 1) tmk_die("tmk_shard: world = %d", world);
    const size_t per = (n + (size_t)world - 1) / (size_t)world;
    return (per + TMK_SHARD_ALIGN - 1) / TMK_SHARD_ALIGN * TMK_SHARD_ALIGN;
}
extern "C" void tmk_shard_bounds(size_t n, int world, int rank, size_t *lo, size_t *hi) {
    if (rank < 0 || rank >= world) tmk_die("tmk_shard: rank %d outside 0..%d", rank, world - 1);
    const size_t per = tmk_shard_per_rank(n, world);
    size_t a = per * (size_t)rank;
    if (a > n) a = n;
    size_t b = a + per;
    if (b > n) b = n;
    *lo = a;
    *hi = b;
}
extern "C" size_t tmk_shard_words(size_t n, int world) { return tmk_shard_per_rank(n, world) / 32; }

/* zero the words of this rank's slot that no unit of its slice covers (a short or empty last slice) */
static void zero_slot_tail(uint32_t *slot, size_t n_slice, size_t wpr, cudaStream_t s) {
    const size_t used = (n_slice + 31) / 32;
    if (used < wpr) CK(cudaMemsetAsync(slot + used, 0, sizeof(uint32_t) * (wpr - used), s));
}

/* ------------------------------------------------------------------ one process per GPU */
struct tmk_shard_comm {
    ncclComm_t comm;
    int rank, world, dev;
    int overlap;              /* 1: the all-gather runs on `cs`, beside the kernels of the next call */
    cudaStream_t cs;          /* high priority: tiny, and every peer waits for it */
    cudaEvent_t ev_done;
    /* gathers in flight on `cs` (overlap mode), by the bitmap they complete: a later call that writes the same
     * bitmap first makes its stream wait for that gather - and only for that one, so that a caller alternating
     * between two bitmaps overlaps the gather of batch k with the kernels of batch k+1 */
    struct { const void *buf; cudaEvent_t ev; int pending; } fly[2];
    int next_fly;
};

extern "C" void tmk_shard_unique_id(void *id128) {
    ncclUniqueId id;
    NK(ncclGetUniqueId(&id));
    static_assert(sizeof(id) == TMK_SHARD_ID_BYTES, "NCCL unique id size");
    memcpy(id128, &id, sizeof(id));
}

extern "C" struct tmk_shard_comm *tmk_shard_comm_create(int rank, int world, const void *id128) {
    if (world < 1 || rank < 0 || rank >= world) tmk_die("tmk_shard_comm_create: rank %d of %d", rank, world);
    struct tmk_shard_comm *c = new tmk_shard_comm();
    c->rank = rank; c->world = world; c->overlap = 0; c->next_fly = 0;
    CK(cudaGetDevice(&c->dev));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NK(ncclCommInitRank(&c->comm, world, id, rank));
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->cs, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        CK(cudaEventCreateWithFlags(&c->fly[k].ev, cudaEventDisableTiming));
        c->fly[k].buf = nullptr; c->fly[k].pending = 0;
    }
    return c;
}

extern "C" void tmk_shard_comm_destroy(struct tmk_shard_comm *c) {
    if (!c) return;
    cudaStreamSynchronize(c->cs);
    ncclCommDestroy(c->comm);
    cudaEventDestroy(c->ev_done);
    for (int k = 0; k < 2; k++) cudaEventDestroy(c->fly[k].ev);
    cudaStreamDestroy(c->cs);
    delete c;
}
extern "C" int tmk_shard_comm_rank(const struct tmk_shard_comm *c) { return c->rank; }
extern "C" int tmk_shard_comm_world(const struct tmk_shard_comm *c) { return c->world; }
extern "C" void tmk_shard_comm_overlap(struct tmk_shard_comm *c, int enable) { c->overlap = enable ? 1 : 0; }

/* `stream` waits for the gathers still in flight on the communicator's own stream (overlap mode) */
extern "C" void tmk_shard_comm_join(struct tmk_shard_comm *c, void *stream) {
    for (int k = 0; k < 2; k++)
        if (c->fly[k].pending) {
            CK(cudaStreamWaitEvent((cudaStream_t)stream, c->fly[k].ev, 0));
            c->fly[k].pending = 0;
        }
}

/* before `s` writes into d_bits_all again: wait for the gather in flight that still reads / completes it */
static void wait_bitmap_free(struct tmk_shard_comm *c, const void *d_bits_all, cudaStream_t s) {
    for (int k = 0; k < 2; k++)
        if (c->fly[k].pending && c->fly[k].buf == d_bits_all) {
            CK(cudaStreamWaitEvent(s, c->fly[k].ev, 0));
            c->fly[k].pending = 0;
        }
}

/* in place: this rank's words already sit in slot `rank` of d_bits_all */
static void gather_words(struct tmk_shard_comm *c, uint32_t *d_bits_all, size_t wpr, cudaStream_t s) {
    if (c->world == 1) return;
    if (!c->overlap) {
        NK(ncclAllGather(d_bits_all + (size_t)c->rank * wpr, d_bits_all, wpr, ncclUint32, c->comm, s));
        return;
    }
    CK(cudaEventRecord(c->ev_done, s));
    CK(cudaStreamWaitEvent(c->cs, c->ev_done, 0));
    NK(ncclAllGather(d_bits_all + (size_t)c->rank * wpr, d_bits_all, wpr, ncclUint32, c->comm, c->cs));
    const int k = c->next_fly;
    c->next_fly ^= 1;
    if (c->fly[k].pending) CK(cudaStreamWaitEvent(s, c->fly[k].ev, 0)); /* a third bitmap: the oldest gather is joined */
    CK(cudaEventRecord(c->fly[k].ev, c->cs));
    c->fly[k].buf = d_bits_all;
    c->fly[k].pending = 1;
}

extern "C" void aa_rx_cl_check_batch_shard_dev(struct aa_rx_cl *cl, struct tmk_shard_comm *c, size_t n_total, size_t n_sub,
                                               const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                               const void *dQ_slice, int layout, size_t ld, uint32_t *d_bits_all,
                                               void *stream) {
    size_t lo, hi;
    tmk_shard_bounds(n_total, c->world, c->rank, &lo, &hi);
    const size_t wpr = tmk_shard_words(n_total, c->world);
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)tmk_cl_stream(cl);
    uint32_t *slot = d_bits_all + (size_t)c->rank * wpr;
    wait_bitmap_free(c, d_bits_all, s);
    zero_slot_tail(slot, hi - lo, wpr, s);
    aa_rx_cl_check_batch_dev(cl, hi - lo, n_sub, sub_ids, n_q, q_all_base, dQ_slice, layout, ld, slot, s);
    gather_words(c, d_bits_all, wpr, s);
}

extern "C" void aa_rx_cl_check_edges_shard_dev(struct aa_rx_cl *cl, struct tmk_shard_comm *c, size_t n_total, size_t n_sub,
                                               const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                               const void *dQ0_slice, const void *dQ1_slice, int layout, size_t ld,
                                               double seg_len, uint32_t *d_bits_all, int32_t *d_first_slice, void *stream) {
    size_t lo, hi;
    tmk_shard_bounds(n_total, c->world, c->rank, &lo, &hi);
    const size_t wpr = tmk_shard_words(n_total, c->world);
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)tmk_cl_stream(cl);
    uint32_t *slot = d_bits_all + (size_t)c->rank * wpr;
    wait_bitmap_free(c, d_bits_all, s);
    zero_slot_tail(slot, hi - lo, wpr, s);
    aa_rx_cl_check_edges_dev(cl, hi - lo, n_sub, sub_ids, n_q, q_all_base, dQ0_slice, dQ1_slice, layout, ld, seg_len, slot,
                             d_first_slice, s);
    gather_words(c, d_bits_all, wpr, s);
}

/* ------------------------------------------------------------------ one process, all devices */
struct aa_rx_cl_multi {
    const struct aa_rx_sg *sg;
    std::vector<int> dev;
    std::vector<struct aa_rx_cl *> cl;
    std::vector<ncclComm_t> comm; /* built by the first device-resident call */
    size_t n_invalid;
};

struct DevGuard { /* calls leave the caller's current device as they found it */
    int cur;
    DevGuard() { cur = 0; cudaGetDevice(&cur); }
    ~DevGuard() { cudaSetDevice(cur); }
};

extern "C" struct aa_rx_cl_multi *aa_rx_cl_multi_create(const struct aa_rx_sg *sg, int n_dev, const int *devices) {
    int visible = 0;
    CK(cudaGetDeviceCount(&visible));
    if (visible < 1) tmk_die("aa_rx_cl_multi_create: no CUDA device (no CPU fallback exists)");
    if (n_dev <= 0) { n_dev = visible; devices = nullptr; }
    DevGuard guard;
    struct aa_rx_cl_multi *m = new aa_rx_cl_multi();
    m->sg = sg;
    m->n_invalid = 0;
    for (int k = 0; k < n_dev; k++) {
        const int d = devices ? devices[k] : k;
        if (d < 0 || d >= visible) tmk_die("aa_rx_cl_multi_create: device %d not among the %d visible devices", d, visible);
        CK(cudaSetDevice(d));
        m->dev.push_back(d);
        m->cl.push_back(aa_rx_cl_create(sg));
    }
    return m;
}

extern "C" void aa_rx_cl_multi_destroy(struct aa_rx_cl_multi *m) {
    if (!m) return;
    DevGuard guard;
    for (size_t k = 0; k < m->cl.size(); k++) {
        cudaSetDevice(m->dev[k]);
        if (!m->comm.empty()) {
            cudaDeviceSynchronize();
            ncclCommDestroy(m->comm[k]);
        }
        aa_rx_cl_destroy(m->cl[k]);
    }
    delete m;
}

extern "C" int aa_rx_cl_multi_device_count(const struct aa_rx_cl_multi *m) { return (int)m->cl.size(); }
extern "C" int aa_rx_cl_multi_device(const struct aa_rx_cl_multi *m, int k) {
    if (k < 0 || (size_t)k >= m->dev.size()) tmk_die("aa_rx_cl_multi_device: index %d", k);
    return m->dev[k];
}
extern "C" struct aa_rx_cl *aa_rx_cl_multi_context(struct aa_rx_cl_multi *m, int k) {
    if (k < 0 || (size_t)k >= m->cl.size()) tmk_die("aa_rx_cl_multi_context: index %d", k);
    return m->cl[k];
}
extern "C" void aa_rx_cl_multi_allow(struct aa_rx_cl_multi *m, aa_rx_frame_id i, aa_rx_frame_id j, int allowed) {
    for (struct aa_rx_cl *c : m->cl) aa_rx_cl_allow(c, i, j, allowed);
}
extern "C" void aa_rx_cl_multi_allow_set(struct aa_rx_cl_multi *m, const struct aa_rx_cl_set *set) {
    for (struct aa_rx_cl *c : m->cl) aa_rx_cl_allow_set(c, set);
}
extern "C" void aa_rx_cl_multi_allow_name(struct aa_rx_cl_multi *m, const char *f0, const char *f1, int allowed) {
    for (struct aa_rx_cl *c : m->cl) aa_rx_cl_allow_name(c, f0, f1, allowed);
}
extern "C" void aa_rx_cl_multi_track_collisions(struct aa_rx_cl_multi *m, int enable) {
    for (struct aa_rx_cl *c : m->cl) aa_rx_cl_track_collisions(c, enable);
}
extern "C" void aa_rx_cl_multi_batch_collisions(struct aa_rx_cl_multi *m, struct aa_rx_cl_set *set) {
    for (struct aa_rx_cl *c : m->cl) aa_rx_cl_batch_collisions(c, set); /* ORs into the set */
}

static bool host_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

/* run f(k) for every device with a non-empty slice: in the calling thread when the rows are pinned (the copies
 * are asynchronous and all devices get going at once), one thread per device when they are pageable (the
 * driver stages pageable rows through its own pinned buffer inside cudaMemcpyAsync - one thread would feed the
 * devices one after the other) */
template <class F> static void for_devices(struct aa_rx_cl_multi *m, bool threads, F f) {
    const size_t n = m->cl.size();
    if (!threads || n == 1) {
        for (size_t k = 0; k < n; k++) { CK(cudaSetDevice(m->dev[k])); f(k); }
        return;
    }
    std::vector<std::thread> th;
    for (size_t k = 0; k < n; k++)
        th.emplace_back([&, k] { CK(cudaSetDevice(m->dev[k])); f(k); });
    for (auto &t : th) t.join();
}

static size_t multi_host(struct aa_rx_cl_multi *m, size_t n, size_t n_sub, const aa_rx_config_id *sub_ids, size_t n_q,
                         const double *q_base, const void *Q0, const void *Q1, int layout, size_t ldQ, double seg_len,
                         uint32_t *valid_bits, int32_t *first_invalid, bool edges) {
    DevGuard guard;
    const int world = (int)m->cl.size();
    const size_t esz = layout == TMK_Q_ROWS_F64 ? 8 : 4;
    const bool threads = n >= (size_t)65536 * world && !host_pinned(Q0);
    for_devices(m, threads, [&](size_t k) {
        size_t lo, hi;
        tmk_shard_bounds(n, world, (int)k, &lo, &hi);
        if (hi == lo) return;
        const char *q0 = (const char *)Q0 + lo * ldQ * esz;
        if (edges)
            tmk_cl_edges_host_begin(m->cl[k], hi - lo, n_sub, sub_ids, n_q, q_base, (const double *)q0,
                                    (const double *)((const char *)Q1 + lo * ldQ * esz), ldQ, seg_len, valid_bits + lo / 32,
                                    first_invalid ? first_invalid + lo : nullptr);
        else
            tmk_cl_batch_host_begin(m->cl[k], hi - lo, n_sub, sub_ids, n_q, q_base, q0, layout, ldQ, valid_bits + lo / 32);
    });
    size_t bad = 0;
    for (int k = 0; k < world; k++) {
        size_t lo, hi;
        tmk_shard_bounds(n, world, k, &lo, &hi);
        if (hi == lo) continue;
        CK(cudaSetDevice(m->dev[k]));
        bad += tmk_cl_host_end(m->cl[k], hi - lo, valid_bits + lo / 32);
    }
    m->n_invalid = bad;
    return bad;
}

extern "C" size_t aa_rx_cl_check_batch_multi(struct aa_rx_cl_multi *m, size_t n_cfg, size_t n_sub,
                                             const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                             const double *Q_sub, size_t ldQ, uint32_t *valid_bits) {
    return multi_host(m, n_cfg, n_sub, sub_ids, n_q, q_all_base, Q_sub, nullptr, TMK_Q_ROWS_F64, ldQ, 0.0, valid_bits, nullptr,
                      false);
}
extern "C" size_t aa_rx_cl_check_batch_multi_f32(struct aa_rx_cl_multi *m, size_t n_cfg, size_t n_sub,
                                                 const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                                 const float *Q_sub, size_t ldQ, uint32_t *valid_bits) {
    return multi_host(m, n_cfg, n_sub, sub_ids, n_q, q_all_base, Q_sub, nullptr, TMK_Q_ROWS_F32, ldQ, 0.0, valid_bits, nullptr,
                      false);
}
extern "C" size_t aa_rx_cl_check_edges_multi(struct aa_rx_cl_multi *m, size_t n_edge, size_t n_sub,
                                             const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                             const double *Q0, const double *Q1, size_t ldQ, double seg_len,
                                             uint32_t *valid_bits, int32_t *first_invalid) {
    return multi_host(m, n_edge, n_sub, sub_ids, n_q, q_all_base, Q0, Q1, TMK_Q_ROWS_F64, ldQ, seg_len, valid_bits,
                      first_invalid, true);
}

static void multi_comms(struct aa_rx_cl_multi *m) {
    if (!m->comm.empty() || m->cl.size() == 1) return;
    /* a device may be listed twice for the host-buffer calls (two contexts, two slices: what the single-GPU
     * tests do); NCCL needs one rank per device */
    for (size_t k = 0; k < m->dev.size(); k++)
        for (size_t j = 0; j < k; j++)
            if (m->dev[j] == m->dev[k]) tmk_die("aa_rx_cl_multi: device %d listed twice - the device-resident calls need distinct devices", m->dev[k]);
    m->comm.resize(m->cl.size());
    NK(ncclCommInitAll(m->comm.data(), (int)m->dev.size(), m->dev.data()));
}

/* device-resident: dQ[k] is device k's slice (rows lo_k .. hi_k of the batch, in `layout` with leading dimension
 * ld), d_bits_all[k] device k's copy of the global bitmap (n_dev * tmk_shard_words(n, n_dev) words).  Enqueued on
 * every context's own stream and not synchronised (aa_rx_cl_multi_sync does that); afterwards every device
 * holds the whole bitmap. */
extern "C" void aa_rx_cl_check_batch_multi_dev(struct aa_rx_cl_multi *m, size_t n_cfg, size_t n_sub,
                                               const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                               const void *const *dQ, int layout, size_t ld, uint32_t *const *d_bits_all) {
    DevGuard guard;
    const int world = (int)m->cl.size();
    const size_t wpr = tmk_shard_words(n_cfg, world);
    multi_comms(m);
    for (int k = 0; k < world; k++) {
        size_t lo, hi;
        tmk_shard_bounds(n_cfg, world, k, &lo, &hi);
        CK(cudaSetDevice(m->dev[k]));
        cudaStream_t s = (cudaStream_t)tmk_cl_stream(m->cl[k]);
        uint32_t *slot = d_bits_all[k] + (size_t)k * wpr;
        zero_slot_tail(slot, hi - lo, wpr, s);
        aa_rx_cl_check_batch_dev(m->cl[k], hi - lo, n_sub, sub_ids, n_q, q_all_base, dQ[k], layout, ld, slot, s);
    }
    if (world == 1) return;
    NK(ncclGroupStart());
    for (int k = 0; k < world; k++)
        NK(ncclAllGather(d_bits_all[k] + (size_t)k * wpr, d_bits_all[k], wpr, ncclUint32, m->comm[k],
                         (cudaStream_t)tmk_cl_stream(m->cl[k])));
    NK(ncclGroupEnd());
}

extern "C" void aa_rx_cl_check_edges_multi_dev(struct aa_rx_cl_multi *m, size_t n_edge, size_t n_sub,
                                               const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                               const void *const *dQ0, const void *const *dQ1, int layout, size_t ld,
                                               double seg_len, uint32_t *const *d_bits_all, int32_t *const *d_first) {
    DevGuard guard;
    const int world = (int)m->cl.size();
    const size_t wpr = tmk_shard_words(n_edge, world);
    multi_comms(m);
    for (int k = 0; k < world; k++) {
        size_t lo, hi;
        tmk_shard_bounds(n_edge, world, k, &lo, &hi);
        CK(cudaSetDevice(m->dev[k]));
        cudaStream_t s = (cudaStream_t)tmk_cl_stream(m->cl[k]);
        uint32_t *slot = d_bits_all[k] + (size_t)k * wpr;
        zero_slot_tail(slot, hi - lo, wpr, s);
        aa_rx_cl_check_edges_dev(m->cl[k], hi - lo, n_sub, sub_ids, n_q, q_all_base, dQ0[k], dQ1[k], layout, ld, seg_len, slot,
                                 d_first ? d_first[k] : nullptr, s);
    }
    if (world == 1) return;
    NK(ncclGroupStart());
    for (int k = 0; k < world; k++)
        NK(ncclAllGather(d_bits_all[k] + (size_t)k * wpr, d_bits_all[k], wpr, ncclUint32, m->comm[k],
                         (cudaStream_t)tmk_cl_stream(m->cl[k])));
    NK(ncclGroupEnd());
}

extern "C" void aa_rx_cl_multi_sync(struct aa_rx_cl_multi *m) {
    DevGuard guard;
    for (size_t k = 0; k < m->cl.size(); k++) {
        CK(cudaSetDevice(m->dev[k]));
        aa_rx_cl_sync(m->cl[k], nullptr);
    }
}

/* ------------------------------------------------------------------ pinned host buffers for the callers
 * of the host-buffer entry points (pageable rows are staged by the driver at a fraction of the PCIe rate) */
extern "C" void *aa_rx_host_alloc(size_t bytes) {
    void *p = nullptr;
    CK(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable));
    return p;
}
extern "C" void aa_rx_host_free(void *p) {
    if (p) CK(cudaFreeHost(p));
}
extern "C" int aa_rx_host_register(void *p, size_t bytes) {
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}
extern "C" void aa_rx_host_unregister(void *p) {
    if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
}
