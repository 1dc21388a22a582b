/* cl.cu - collision context: device tables, batch images, kernel launches, the C ABI on top.
 *
 * Reference call shapes: demo/tmpexec.c:397-421 (aa_rx_cl_create / aa_rx_cl_check /
 * aa_rx_cl_set_get), :118 (allow_config), :322-325 (aa_rx_sg_tf).  Host code is C-style; it is a
 * .cu only because it launches kernels.  No CPU fallback: every compute entry point needs a
 * CUDA device and aborts loudly without one.
 */
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <chrono>
#include <mutex>
#include <vector>

#include "kernels.cuh"
#include "pipeline.cuh"
#include "wave.cuh"
#include "ik.cuh"
#include "tmk_internal.h"

using namespace tmk;

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            tmk_die("CUDA error %s at %s:%d: %s (no CPU fallback exists)", cudaGetErrorName(e_), __FILE__, \
                    __LINE__, cudaGetErrorString(e_));                                               \
    } while (0)

/* host-side spans to stderr when TMK_PROFILE_HOST is set (tuning aid) */
struct HostSpan {
    static bool on() { static int v = -1; if (v < 0) v = getenv("TMK_PROFILE_HOST") ? 1 : 0; return v == 1; }
    std::chrono::steady_clock::time_point t0;
    const char *what;
    explicit HostSpan(const char *w) : what(w) { if (on()) t0 = std::chrono::steady_clock::now(); }
    void mark(const char *label) {
        if (!on()) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "tmkcl host: %-18s %-22s %8.1f us\n", what, label, std::chrono::duration<double, std::micro>(t1 - t0).count());
        t0 = t1;
    }
};

/* ------------------------------------------------------------------ small device buffer helper */
/* Device buffers and pinned slots are recycled through a process-wide pool: a task-motion planner builds a
 * new collision context for every motion query (start-state allowance, reparented objects), and
 * cudaMalloc / cudaMallocHost / cudaFree cost 1-3 ms per context - more than the query's validity checks. */
struct DevPool {
    struct Ent { void *p; size_t cap; int dev; };
    std::mutex mu;
    std::vector<Ent> free_;
    size_t held = 0;
    static constexpr size_t HELD_MAX = size_t(4) << 30;
    void *take(size_t bytes, size_t *cap) {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = free_.size();
            for (size_t k = 0; k < free_.size(); k++)
                if (free_[k].dev == dev && free_[k].cap >= bytes && free_[k].cap <= 4 * bytes + 4096 &&
                    (best == free_.size() || free_[k].cap < free_[best].cap))
                    best = k;
            if (best != free_.size()) {
                Ent e = free_[best];
                free_[best] = free_.back();
                free_.pop_back();
                held -= e.cap;
                *cap = e.cap;
                return e.p;
            }
        }
        void *p = nullptr;
        *cap = bytes + bytes / 4 + 256;
        CK(cudaMalloc(&p, *cap));
        return p;
    }
    /* the caller guarantees that no work in flight still touches p */
    void give(void *p, size_t cap) {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> g(mu);
        if (held + cap > HELD_MAX) { cudaFree(p); return; }
        free_.push_back({p, cap, dev});
        held += cap;
    }
};
static DevPool g_dev_pool;

struct PinnedPool { /* 64-byte pinned slots (the counters a batch call reads back) */
    std::mutex mu;
    std::vector<void *> free_;
    void *take() {
        std::lock_guard<std::mutex> g(mu);
        if (free_.empty()) {
            char *blk = nullptr;
            CK(cudaMallocHost((void **)&blk, 4096));
            for (int k = 0; k < 64; k++) free_.push_back(blk + 64 * k);
        }
        void *p = free_.back();
        free_.pop_back();
        return p;
    }
    void give(void *p) { std::lock_guard<std::mutex> g(mu); free_.push_back(p); }
};
static PinnedPool g_pinned_pool;

struct StreamSet { cudaStream_t stream, copy_stream, aux_stream; cudaEvent_t ev[8], ev_aux[4]; int dev; };
struct StreamPool { /* the two streams and the copy events of a collision context (idle when returned) */
    std::mutex mu;
    std::vector<StreamSet> free_;
    StreamSet take() {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t k = 0; k < free_.size(); k++)
                if (free_[k].dev == dev) {
                    StreamSet ss = free_[k];
                    free_[k] = free_.back();
                    free_.pop_back();
                    return ss;
                }
        }
        StreamSet ss;
        ss.dev = dev;
        CK(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ss.copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ss.aux_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 8; k++) CK(cudaEventCreateWithFlags(&ss.ev[k], cudaEventDisableTiming));
        for (int k = 0; k < 4; k++) CK(cudaEventCreateWithFlags(&ss.ev_aux[k], cudaEventDisableTiming));
        return ss;
    }
    void give(const StreamSet &ss) { std::lock_guard<std::mutex> g(mu); free_.push_back(ss); }
};
static StreamPool g_stream_pool;

struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    void need(size_t bytes) {
        if (bytes <= cap) return;
        if (p) { /* growth: earlier launches may still read the old buffer */
            CK(cudaDeviceSynchronize());
            g_dev_pool.give(p, cap);
        }
        p = g_dev_pool.take(bytes, &cap);
    }
    /* only after the streams that used the buffer were synchronised */
    void release() {
        if (p) g_dev_pool.give(p, cap);
        p = nullptr;
        cap = 0;
    }
};

template <typename T> static void upload(DBuf &b, const std::vector<T> &v, cudaStream_t s) {
    b.need(std::max<size_t>(16, v.size() * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

/* ------------------------------------------------------------------ scene-graph device mirror (FK) */
struct FkImage { /* cached per scene graph and device: rebuilt when the arguments change */
    std::vector<int> sub, frames;
    std::vector<double> q_base;
    uint64_t sg_serial = 0;
    bool all_frames = false, valid = false;
    DBuf joints, outs;
    int n_joint = 0, n_out = 0;
};
struct SgDev {
    FkImage *fk = nullptr;     /* tables of aa_rx_sg_tf_batch for the last (sub_ids, q_base, frames) */
    DBuf frames, q, rel, ab;
    DBuf b_frames, b_q, b_out; /* aa_rx_sg_tf_batch scratch */
    DBuf ik_in, ik_out;        /* tmk_dev_ik: one upload, one launch, one download per goal */
    size_t n_f = 0;
};
/* one mirror per CUDA device that has touched the scene graph (multi-GPU: a replica per device) */
#define TMK_MAX_DEV 16
struct SgDevSet { SgDev *d[TMK_MAX_DEV] = {nullptr}; };

extern "C" void tmk_dev_sg_release(void *dev) {
    SgDevSet *set = (SgDevSet *)dev;
    if (!set) return;
    int cur = 0;
    cudaGetDevice(&cur);
    for (int k = 0; k < TMK_MAX_DEV; k++) {
        SgDev *d = set->d[k];
        if (!d) continue;
        cudaSetDevice(k); /* the buffers go back to the pool under their own device */
        cudaDeviceSynchronize();
        d->frames.release(); d->q.release(); d->rel.release(); d->ab.release();
        d->b_frames.release(); d->b_q.release(); d->b_out.release();
        d->ik_in.release(); d->ik_out.release();
        if (d->fk) { d->fk->joints.release(); d->fk->outs.release(); delete d->fk; }
        delete d;
    }
    cudaSetDevice(cur);
    delete set;
}

static SgDev *sg_dev(const struct aa_rx_sg *sg) {
    struct aa_rx_sg *m = (struct aa_rx_sg *)sg; /* cache slot is logically mutable */
    int cur = 0;
    CK(cudaGetDevice(&cur));
    if (cur < 0 || cur >= TMK_MAX_DEV) tmk_die("device ordinal %d outside 0..%d", cur, TMK_MAX_DEV - 1);
    static std::mutex mu; /* the per-device contexts of a multi-GPU handle may be built from several threads */
    std::lock_guard<std::mutex> lock(mu);
    if (!m->dev) m->dev = new SgDevSet();
    SgDevSet *set = (SgDevSet *)m->dev;
    if (set->d[cur]) return set->d[cur];
    SgDev *d = new SgDev();
    std::vector<DFrame> fr(sg->n_f);
    for (size_t i = 0; i < sg->n_f; i++) {
        memcpy(fr[i].E, sg->E + 7 * i, sizeof(double) * 7);
        memcpy(fr[i].ax, sg->axis + 3 * i, sizeof(double) * 3);
        fr[i].off = sg->offset[i];
        fr[i].parent = sg->parent[i];
        fr[i].type = sg->type[i];
        fr[i].qidx = sg->qidx[i];
        fr[i].pad = 0;
    }
    upload(d->frames, fr, 0);
    d->n_f = sg->n_f;
    d->q.need(sizeof(double) * (sg->n_q + 1));
    d->rel.need(sizeof(double) * 7 * (sg->n_f + 1));
    d->ab.need(sizeof(double) * 7 * (sg->n_f + 1));
    CK(cudaDeviceSynchronize());
    set->d[cur] = d;
    return d;
}

extern "C" void tmk_dev_fk64(const struct aa_rx_sg *sg, const double *q, double *TF_rel, double *TF_abs) {
    SgDev *d = sg_dev(sg);
    if (sg->n_q) CK(cudaMemcpy(d->q.p, q, sizeof(double) * sg->n_q, cudaMemcpyHostToDevice));
    k_fk_single64<<<1, 32>>>((const DFrame *)d->frames.p, (int)sg->n_f, (const double *)d->q.p, (double *)d->rel.p,
                             (double *)d->ab.p);
    CK(cudaGetLastError());
    CK(cudaMemcpy(TF_rel, d->rel.p, sizeof(double) * 7 * sg->n_f, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(TF_abs, d->ab.p, sizeof(double) * 7 * sg->n_f, cudaMemcpyDeviceToHost));
}

/* Damped-least-squares IK of a kinematic chain for n_seed start guesses, all iterations in one launch
 * (ik.cuh).  frames[0..n_fr) is the parent-linked path from the chain root to the tip.  Q (n_seed x n_sub)
 * holds the seeds on entry and the final iterates on return, err (n_seed x 2) their position / rotation
 * error; a seed converged iff both are below tol.  Returns the number of iterations run. */
extern "C" int tmk_dev_ik(const struct aa_rx_sg *sg, size_t n_fr, const aa_rx_frame_id *frames, size_t n_sub,
                          const aa_rx_config_id *sub_ids, const double *q_all_base, const double *lo, const double *hi,
                          const double *goal, size_t n_seed, double *Q, double *err, int max_iter, int enough,
                          int min_iter, double tol) {
    if (n_seed < 1 || n_seed > 32) tmk_die("tmk_dev_ik: n_seed = %zu outside 1..32", n_seed);
    if (n_sub < 1 || n_sub > TMK_IK_MAXCOL) tmk_die("tmk_dev_ik: n_sub = %zu outside 1..%d", n_sub, TMK_IK_MAXCOL);
    if (n_fr < 1) tmk_die("tmk_dev_ik: empty chain");
    SgDev *d = sg_dev(sg);
    IkParams P;
    memset(&P, 0, sizeof(P));
    P.base[3] = 1.0;
    const int root = sg->parent[frames[0]];
    if (root >= 0) { /* chain hanging off an inner frame: its pose at the base configuration */
        std::vector<double> rel(7 * sg->n_f), ab(7 * sg->n_f);
        tmk_dev_fk64(sg, q_all_base, rel.data(), ab.data());
        memcpy(P.base, &ab[7 * (size_t)root], sizeof(double) * 7);
    }
    memcpy(P.goal, goal, sizeof(double) * 7);
    P.n_fr = (int)n_fr; P.n_sub = (int)n_sub; P.n_seed = (int)n_seed; P.max_iter = max_iter;
    P.enough = enough; P.min_iter = min_iter; P.tol = tol;
    /* one staging block: frames | lo | hi | Q */
    const size_t b_fr = sizeof(IkFrame) * n_fr, b_lim = sizeof(double) * n_sub, b_q = sizeof(double) * n_seed * n_sub;
    std::vector<unsigned char> in(b_fr + 2 * b_lim + b_q);
    IkFrame *fr = (IkFrame *)in.data();
    for (size_t f = 0; f < n_fr; f++) {
        const int i = frames[f];
        if (i < 0 || (size_t)i >= sg->n_f) tmk_die("tmk_dev_ik: bad frame id %d", i);
        if (f > 0 && sg->parent[i] != frames[f - 1]) tmk_die("tmk_dev_ik: frames are not a parent-linked chain");
        memcpy(fr[f].E, sg->E + 7 * (size_t)i, sizeof(double) * 7);
        memcpy(fr[f].ax, sg->axis + 3 * (size_t)i, sizeof(double) * 3);
        fr[f].off = sg->offset[i];
        fr[f].type = sg->type[i];
        fr[f].col = -1;
        fr[f].qfix = 0.0;
        if (fr[f].type != TMK_FIXED) {
            fr[f].qfix = q_all_base[sg->qidx[i]];
            for (size_t k = 0; k < n_sub; k++)
                if (sub_ids[k] == sg->qidx[i]) fr[f].col = (int)k;
        }
    }
    memcpy(in.data() + b_fr, lo, b_lim);
    memcpy(in.data() + b_fr + b_lim, hi, b_lim);
    memcpy(in.data() + b_fr + 2 * b_lim, Q, b_q);
    d->ik_in.need(in.size());
    const size_t b_err = sizeof(double) * 2 * n_seed;
    d->ik_out.need(b_err + 16);
    CK(cudaMemcpy(d->ik_in.p, in.data(), in.size(), cudaMemcpyHostToDevice));
    unsigned char *di = (unsigned char *)d->ik_in.p, *dout = (unsigned char *)d->ik_out.p;
    k_ik_dls<<<1, 32>>>((const IkFrame *)di, P, (const double *)(di + b_fr), (const double *)(di + b_fr + b_lim),
                        (double *)(di + b_fr + 2 * b_lim), (double *)dout, (int *)(dout + b_err));
    CK(cudaGetLastError());
    std::vector<unsigned char> out(b_err + 16);
    CK(cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Q, di + b_fr + 2 * b_lim, b_q, cudaMemcpyDeviceToHost));
    memcpy(err, out.data(), b_err);
    int iters = 0;
    memcpy(&iters, out.data() + b_err, sizeof(int));
    return iters;
}

/* ------------------------------------------------------------------ mesh BVH (host build) */
struct HostMesh {
    std::vector<BvhNode> nodes;
    std::vector<double> tris; /* 9 per triangle, reordered by the tree */
    double centre[3], brad;
};

static int bvh_build(HostMesh &hm, std::vector<int> &ids, const std::vector<double> &src, int lo, int hi,
                     std::vector<double> &out_tris) {
    int me = (int)hm.nodes.size();
    hm.nodes.push_back(BvhNode());
    double bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
    double cl[3] = {INFINITY, INFINITY, INFINITY}, ch[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = lo; k < hi; k++) {
        const double *t = &src[9 * (size_t)ids[k]];
        for (int a = 0; a < 3; a++) {
            double c = (t[a] + t[3 + a] + t[6 + a]) / 3;
            cl[a] = std::min(cl[a], c); ch[a] = std::max(ch[a], c);
            for (int v = 0; v < 3; v++) { bl[a] = std::min(bl[a], t[3 * v + a]); bh[a] = std::max(bh[a], t[3 * v + a]); }
        }
    }
    BvhNode nd;
    for (int a = 0; a < 3; a++) {
        /* round outward so the float box contains the double one */
        nd.lo[a] = nextafterf((float)bl[a], -INFINITY);
        nd.hi[a] = nextafterf((float)bh[a], INFINITY);
    }
    if (hi - lo <= 2) {
        nd.a = (int)(out_tris.size() / 9);
        nd.b = hi - lo;
        for (int k = lo; k < hi; k++) out_tris.insert(out_tris.end(), &src[9 * (size_t)ids[k]], &src[9 * (size_t)ids[k]] + 9);
        hm.nodes[me] = nd;
        return me;
    }
    int ax = 0;
    if (ch[1] - cl[1] > ch[ax] - cl[ax]) ax = 1;
    if (ch[2] - cl[2] > ch[ax] - cl[ax]) ax = 2;
    int mid = (lo + hi) / 2;
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int x, int y) {
        const double *tx = &src[9 * (size_t)x], *ty = &src[9 * (size_t)y];
        return tx[ax] + tx[3 + ax] + tx[6 + ax] < ty[ax] + ty[3 + ax] + ty[6 + ax];
    });
    int l = bvh_build(hm, ids, src, lo, mid, out_tris);
    int r = bvh_build(hm, ids, src, mid, hi, out_tris);
    nd.a = l;
    nd.b = -r - 1;
    hm.nodes[me] = nd;
    return me;
}

static void mesh_compile(const struct aa_rx_mesh *m, HostMesh &hm) {
    std::vector<double> src(9 * m->nt);
    for (size_t t = 0; t < m->nt; t++)
        for (int v = 0; v < 3; v++)
            for (int a = 0; a < 3; a++) src[9 * t + 3 * v + a] = (double)m->verts[3 * (size_t)m->idx[3 * t + v] + a];
    std::vector<int> ids(m->nt);
    for (size_t t = 0; t < m->nt; t++) ids[t] = (int)t;
    hm.nodes.clear();
    hm.tris.clear();
    bvh_build(hm, ids, src, 0, (int)m->nt, hm.tris);
    /* bounding sphere: centre of the AABB, radius = farthest vertex */
    const BvhNode &root = hm.nodes[0];
    for (int a = 0; a < 3; a++) hm.centre[a] = 0.5 * ((double)root.lo[a] + (double)root.hi[a]);
    double r2 = 0;
    for (size_t i = 0; i < hm.tris.size() / 3; i++) {
        double d2 = 0;
        for (int a = 0; a < 3; a++) d2 += (hm.tris[3 * i + a] - hm.centre[a]) * (hm.tris[3 * i + a] - hm.centre[a]);
        r2 = std::max(r2, d2);
    }
    hm.brad = sqrt(r2) * (1 + 1e-12);
}

static std::mutex g_mesh_mu;
extern "C" void tmk_mesh_compiled_release(void *compiled) { delete (HostMesh *)compiled; }

/* ------------------------------------------------------------------ the collision context */
struct ImageDev {
    DBuf jf, jd, mgf, mgd, sgf, sgd, pairs, sq, jg, pa, ps4, tpairs, tkeys, ab4, paabb, g_cells, g_items, g_live, g_c4, g_rstart, g_rc, g_rid;
    Image<float> f;
    Image<double> d;
    std::vector<uint32_t> h_pairs;
    std::vector<int> pair_fa, pair_fb; /* frames of each pair (for collision tracking) */
    std::vector<int> sub;
    std::vector<double> q_base;
    uint64_t allow_serial = 0;
    bool valid = false;
    uint64_t serial = 0; /* bumped whenever the image is rebuilt: cached CUDA graphs refer to its buffers */
    TileCfg wcfg, wcfg_e, wcfg_flat;
};

struct aa_rx_cl {
    const struct aa_rx_sg *sg;
    int dev; /* the CUDA device the context lives on: every call on it must be made with that device current */
    size_t n_f, n_q, n_g, wpr;
    uint32_t *allowed; /* own copy */
    uint64_t allow_serial;
    cudaStream_t stream;
    cudaStream_t copy_stream; /* host-buffer entry points: H2D of chunk k+1 overlaps the kernels of chunk k */
    cudaStream_t aux_stream;  /* the deferred narrowphase forks here: convex items beside the mesh traversal */
    cudaEvent_t ev_copy[8], ev_aux[4];

    /* frame-relative geometry + meshes */
    std::vector<DGeom<double>> geoms;
    std::vector<BvhNode> mesh_root; /* root box of every mesh (the broadphase sees a static mesh as this box) */
    DBuf d_geoms, d_nodes, d_trisf, d_trisd, d_meshf, d_meshd;

    /* generic pair list for aa_rx_cl_check */
    std::vector<uint2> gpairs;
    DBuf d_gpairs, d_gout, d_tf;
    uint64_t gpairs_serial;

    ImageDev img;

    /* batch scratch */
    DBuf d_spill, d_mesh, d_conv, d_tri;
    DBuf d_wpose, d_whit, d_wqueue, d_wpend; /* wavefront pipeline of large scenes (wave.cuh) */
    DBuf d_q, d_q1, d_bits, d_unc, d_cnt, d_first, d_pairhit, d_stats, d_nd, d_alive0, d_alive1, d_efirst;
    uint32_t *h_cnt; /* pinned */
    int track, stats_on;
    std::vector<uint8_t> last_pairhit;
    struct tmk_batch_stats stats;
    int use_tile; /* 1: tile-pipeline kernel (default), 0: thread-per-configuration kernel */
    /* the swept-edge pipeline is ~45 short launches per batch: captured once per (arguments, buffers) into
     * a CUDA graph and replayed, otherwise the host cannot enqueue as fast as the GPU executes */
    struct EdgeGraph { uint64_t key[12]; cudaGraphExec_t exec; uint64_t used; size_t n_kernels; };
    std::vector<EdgeGraph> edge_graphs;
    std::vector<std::array<uint64_t, 12>> edge_seen; /* argument keys seen once (no graph yet) */
    uint64_t graph_clock = 0;
    UncList cur_unc;      /* undecided list of the batch in flight */
    cudaEvent_t *cur_ev;  /* its timing events (NULL when timing is off) */

    /* device-time instrumentation: events around the main kernel / the re-check kernel of every launch */
    int timing;
    std::vector<cudaEvent_t> ev_pool; /* triples: before main, after main, after re-check */
    size_t ev_used;
};

static cudaEvent_t *timing_begin(struct aa_rx_cl *cl, cudaStream_t s) {
    if (!cl->timing) return nullptr;
    if (cl->ev_used + 3 > cl->ev_pool.size()) {
        size_t old = cl->ev_pool.size();
        cl->ev_pool.resize(old + 96);
        for (size_t k = old; k < cl->ev_pool.size(); k++) CK(cudaEventCreate(&cl->ev_pool[k]));
    }
    cudaEvent_t *e = &cl->ev_pool[cl->ev_used];
    cl->ev_used += 3;
    CK(cudaEventRecord(e[0], s));
    return e;
}

static void build_geoms(struct aa_rx_cl *cl) {
    const struct aa_rx_sg *sg = cl->sg;
    HostSpan hs("build_geoms");
    /* the BVH of a mesh is built once and kept with the mesh handle: scene-graph copies share it */
    std::vector<const HostMesh *> hmesh(sg->n_mesh);
    cl->mesh_root.resize(sg->n_mesh);
    {
        std::lock_guard<std::mutex> g(g_mesh_mu);
        for (size_t m = 0; m < sg->n_mesh; m++) {
            struct aa_rx_mesh *mesh = sg->meshes[m];
            if (!mesh->compiled) {
                HostMesh *hm = new HostMesh();
                mesh_compile(mesh, *hm);
                /* k_mesh_tasks packs (item slot, node or triangle index) into one word: 24 bits of index */
                if (hm->nodes.size() >= (size_t(1) << 24) || hm->tris.size() / 9 >= (size_t(1) << 24)) tmk_die("mesh: more than 2^24 triangles");
                mesh->compiled = hm;
            }
            hmesh[m] = (const HostMesh *)mesh->compiled;
            cl->mesh_root[m] = hmesh[m]->nodes[0];
        }
    }
    hs.mark("mesh_compile");
    cl->geoms.resize(sg->n_g);
    for (size_t g = 0; g < sg->n_g; g++) {
        DGeom<double> &G = cl->geoms[g];
        memset(&G, 0, sizeof(G));
        G.tf[3] = 1;
        G.shape = sg->g_shape[g];
        G.frame = sg->g_frame[g];
        G.slot = -1;
        G.mesh = sg->g_mesh[g];
        const double *dim = sg->g_dim + 3 * g;
        switch (G.shape) {
        case TMK_SPHERE: G.dim[0] = dim[0]; G.brad = dim[0]; break;
        case TMK_BOX:
            for (int a = 0; a < 3; a++) G.dim[a] = 0.5 * dim[a];
            G.brad = sqrt(G.dim[0] * G.dim[0] + G.dim[1] * G.dim[1] + G.dim[2] * G.dim[2]);
            break;
        case TMK_CYLINDER: /* base at origin, +z: centred core shifted by h/2 */
            G.dim[0] = dim[1]; G.dim[1] = 0.5 * dim[0];
            G.tf[6] = 0.5 * dim[0];
            G.brad = sqrt(G.dim[0] * G.dim[0] + G.dim[1] * G.dim[1]);
            break;
        default:
            for (int a = 0; a < 3; a++) G.dim[a] = hmesh[G.mesh]->centre[a];
            G.brad = hmesh[G.mesh]->brad;
        }
        G.brad *= (1 + 1e-12);
    }
    /* meshes -> device */
    std::vector<BvhNode> nodes;
    std::vector<double> trisd;
    std::vector<float> trisf;
    std::vector<size_t> noff, toff;
    for (const HostMesh *hm : hmesh) {
        noff.push_back(nodes.size());
        toff.push_back(trisd.size());
        nodes.insert(nodes.end(), hm->nodes.begin(), hm->nodes.end());
        for (size_t t = 0; t < hm->tris.size() / 9; t++) { /* 9 -> TMK_TRI_STRIDE values per triangle */
            trisd.insert(trisd.end(), hm->tris.begin() + 9 * t, hm->tris.begin() + 9 * t + 9);
            trisd.insert(trisd.end(), TMK_TRI_STRIDE - 9, 0.0);
        }
    }
    trisf.assign(trisd.begin(), trisd.end());
    upload(cl->d_nodes, nodes, cl->stream);
    upload(cl->d_trisf, trisf, cl->stream);
    upload(cl->d_trisd, trisd, cl->stream);
    std::vector<MeshRef<float>> mf(hmesh.size());
    std::vector<MeshRef<double>> md(hmesh.size());
    for (size_t m = 0; m < hmesh.size(); m++) {
        mf[m].nodes = md[m].nodes = (const BvhNode *)cl->d_nodes.p + noff[m];
        mf[m].tris = (const float *)cl->d_trisf.p + toff[m];
        md[m].tris = (const double *)cl->d_trisd.p + toff[m];
        md[m].brad = hmesh[m]->brad; mf[m].brad = (float)hmesh[m]->brad;
        md[m].cx = hmesh[m]->centre[0]; md[m].cy = hmesh[m]->centre[1]; md[m].cz = hmesh[m]->centre[2];
        mf[m].cx = (float)md[m].cx; mf[m].cy = (float)md[m].cy; mf[m].cz = (float)md[m].cz;
    }
    upload(cl->d_meshf, mf, cl->stream);
    upload(cl->d_meshd, md, cl->stream);
    upload(cl->d_geoms, cl->geoms, cl->stream);
    CK(cudaStreamSynchronize(cl->stream));
    hs.mark("uploads");
}

extern "C" struct aa_rx_cl *aa_rx_cl_create(const struct aa_rx_sg *sg) {
    if (!sg || !sg->inited) tmk_die("aa_rx_cl_create: scene graph not initialised");
    if (!sg->cl_inited) tmk_die("aa_rx_cl_create: call aa_rx_sg_cl_init first");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 1) tmk_die("aa_rx_cl_create: no CUDA device (no CPU fallback exists)");
    HostSpan hs("cl_create");
    struct aa_rx_cl *cl = new aa_rx_cl();
    cl->sg = sg;
    CK(cudaGetDevice(&cl->dev));
    cl->n_f = sg->n_f; cl->n_q = sg->n_q; cl->n_g = sg->n_g; cl->wpr = sg->wpr;
    cl->allowed = (uint32_t *)malloc(sizeof(uint32_t) * (cl->n_f * cl->wpr + 1));
    memcpy(cl->allowed, sg->allowed, sizeof(uint32_t) * cl->n_f * cl->wpr);
    cl->allow_serial = 1;
    cl->gpairs_serial = 0;
    cl->track = 0; cl->stats_on = 0;
    cl->timing = 0; cl->ev_used = 0;
    cl->use_tile = getenv("TMK_NO_TILE") ? 0 : 1;
    memset(&cl->stats, 0, sizeof(cl->stats));
    {
        StreamSet ss = g_stream_pool.take();
        cl->stream = ss.stream; cl->copy_stream = ss.copy_stream; cl->aux_stream = ss.aux_stream;
        memcpy(cl->ev_copy, ss.ev, sizeof(ss.ev));
        memcpy(cl->ev_aux, ss.ev_aux, sizeof(ss.ev_aux));
    }
    hs.mark("streams+events");
    cl->h_cnt = (uint32_t *)g_pinned_pool.take();
    hs.mark("pinned");
    build_geoms(cl);
    hs.mark("build_geoms");
    return cl;
}

extern "C" void aa_rx_cl_destroy(struct aa_rx_cl *cl) {
    if (!cl) return;
    HostSpan hs("cl_destroy");
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (cur_dev != cl->dev) cudaSetDevice(cl->dev); /* buffers and streams go back to the pools of their device */
    cudaDeviceSynchronize(); /* the scratch buffers go back to the pool: nothing may still use them */
    DBuf *bufs[] = {&cl->d_geoms, &cl->d_nodes, &cl->d_trisf, &cl->d_trisd, &cl->d_meshf, &cl->d_meshd, &cl->d_gpairs,
                    &cl->d_gout, &cl->d_tf, &cl->d_spill, &cl->d_mesh, &cl->d_conv, &cl->d_tri, &cl->d_wpose, &cl->d_whit, &cl->d_wqueue, &cl->d_wpend, &cl->d_q, &cl->d_q1, &cl->d_bits, &cl->d_unc, &cl->d_cnt, &cl->d_first,
                    &cl->d_pairhit, &cl->d_stats, &cl->d_nd, &cl->d_alive0, &cl->d_alive1, &cl->d_efirst, &cl->img.jf, &cl->img.jd, &cl->img.mgf, &cl->img.mgd, &cl->img.sgf,
                    &cl->img.sgd, &cl->img.pairs, &cl->img.sq, &cl->img.jg, &cl->img.pa, &cl->img.ps4, &cl->img.tpairs, &cl->img.tkeys, &cl->img.ab4, &cl->img.paabb, &cl->img.g_cells, &cl->img.g_items,
                    &cl->img.g_live, &cl->img.g_c4, &cl->img.g_rstart, &cl->img.g_rc, &cl->img.g_rid};
    for (DBuf *b : bufs) b->release();
    for (cudaEvent_t e : cl->ev_pool) cudaEventDestroy(e);
    for (auto &g : cl->edge_graphs) cudaGraphExecDestroy(g.exec);
    g_pinned_pool.give(cl->h_cnt);
    {
        StreamSet ss;
        ss.stream = cl->stream; ss.copy_stream = cl->copy_stream; ss.aux_stream = cl->aux_stream;
        memcpy(ss.ev, cl->ev_copy, sizeof(ss.ev));
        memcpy(ss.ev_aux, cl->ev_aux, sizeof(ss.ev_aux));
        CK(cudaGetDevice(&ss.dev));
        g_stream_pool.give(ss);
    }
    free(cl->allowed);
    if (cur_dev != cl->dev) cudaSetDevice(cur_dev);
    delete cl;
    hs.mark("all");
}

extern "C" void aa_rx_cl_allow(struct aa_rx_cl *cl, aa_rx_frame_id i, aa_rx_frame_id j, int allowed) {
    if (i < 0 || j < 0 || (size_t)i >= cl->n_f || (size_t)j >= cl->n_f) tmk_die("aa_rx_cl_allow: bad frame id");
    tmk_bit_put(cl->allowed, cl->wpr, (size_t)i, (size_t)j, allowed);
    tmk_bit_put(cl->allowed, cl->wpr, (size_t)j, (size_t)i, allowed);
    cl->allow_serial++;
}
extern "C" void aa_rx_cl_allow_set(struct aa_rx_cl *cl, const struct aa_rx_cl_set *set) {
    if (set->n_f != cl->n_f) tmk_die("aa_rx_cl_allow_set: size mismatch");
    for (size_t k = 0; k < cl->n_f * cl->wpr; k++) cl->allowed[k] |= set->bits[k];
    cl->allow_serial++;
}
extern "C" void aa_rx_cl_allow_name(struct aa_rx_cl *cl, const char *f0, const char *f1, int allowed) {
    int i = aa_rx_sg_frame_id(cl->sg, f0), j = aa_rx_sg_frame_id(cl->sg, f1);
    if (i < 0 || j < 0) tmk_die("aa_rx_cl_allow_name: unknown frame `%s' or `%s'", f0, f1);
    aa_rx_cl_allow(cl, i, j, allowed);
}

static inline bool pair_live(const struct aa_rx_cl *cl, int fa, int fb) {
    return fa != fb && !tmk_bit_get(cl->allowed, cl->wpr, (size_t)fa, (size_t)fb);
}

/* ------------------------------------------------------------------ aa_rx_cl_check */
static void run_pairs_tf(struct aa_rx_cl *cl, const double *tf_dense, const std::vector<uint2> &pairs, DBuf &d_pairs,
                         std::vector<uint8_t> &out) {
    out.assign(pairs.size(), 0);
    if (pairs.empty()) return;
    cl->d_tf.need(sizeof(double) * 7 * cl->n_f);
    CK(cudaMemcpyAsync(cl->d_tf.p, tf_dense, sizeof(double) * 7 * cl->n_f, cudaMemcpyHostToDevice, cl->stream));
    cl->d_gout.need(pairs.size());
    int n = (int)pairs.size();
    k_pairs_tf64<<<(n + 63) / 64, 64, 0, cl->stream>>>((const double *)cl->d_tf.p, (const DGeom<double> *)cl->d_geoms.p,
                                                        (const uint2 *)d_pairs.p, n,
                                                        (const MeshRef<double> *)cl->d_meshd.p, (uint8_t *)cl->d_gout.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out.data(), cl->d_gout.p, pairs.size(), cudaMemcpyDeviceToHost, cl->stream));
    CK(cudaStreamSynchronize(cl->stream));
}

extern "C" int aa_rx_cl_check(struct aa_rx_cl *cl, size_t n_tf, const double *TF, size_t ldTF,
                              struct aa_rx_cl_set *set) {
    if (n_tf < cl->n_f) tmk_die("aa_rx_cl_check: n_tf = %zu < frame count %zu", n_tf, cl->n_f);
    if (ldTF < 7) tmk_die("aa_rx_cl_check: leading dimension < 7");
    if (set && set->n_f != cl->n_f) tmk_die("aa_rx_cl_check: collision set size mismatch");
    if (cl->gpairs_serial != cl->allow_serial) {
        cl->gpairs.clear();
        for (size_t i = 0; i < cl->n_g; i++)
            for (size_t j = i + 1; j < cl->n_g; j++)
                if (pair_live(cl, cl->geoms[i].frame, cl->geoms[j].frame)) cl->gpairs.push_back(make_uint2((unsigned)i, (unsigned)j));
        upload(cl->d_gpairs, cl->gpairs, cl->stream);
        cl->gpairs_serial = cl->allow_serial;
    }
    std::vector<double> dense(7 * cl->n_f);
    for (size_t i = 0; i < cl->n_f; i++) memcpy(&dense[7 * i], TF + i * ldTF, sizeof(double) * 7);
    std::vector<uint8_t> out;
    run_pairs_tf(cl, dense.data(), cl->gpairs, cl->d_gpairs, out);
    int any = 0;
    for (size_t p = 0; p < out.size(); p++)
        if (out[p]) {
            any = 1;
            if (set) aa_rx_cl_set_set(set, cl->geoms[cl->gpairs[p].x].frame, cl->geoms[cl->gpairs[p].y].frame, 1);
        }
    return any;
}

/* ------------------------------------------------------------------ batch image */
template <typename T> static void cvt_joint(const DJoint<double> &s, DJoint<T> &d) {
    for (int k = 0; k < 7; k++) d.E[k] = (T)s.E[k];
    for (int k = 0; k < 3; k++) d.ax[k] = (T)s.ax[k];
    d.off = (T)s.off; d.parent = s.parent; d.type = s.type; d.qslot = s.qslot; d.pad = s.pad;
}
template <typename T> static void cvt_geom(const DGeom<double> &s, DGeom<T> &d) {
    for (int k = 0; k < 7; k++) d.tf[k] = (T)s.tf[k];
    for (int k = 0; k < 3; k++) d.dim[k] = (T)s.dim[k];
    d.brad = (T)(s.brad * (1 + 1e-6)); /* bounding radii only ever round up */
    d.shape = s.shape; d.slot = s.slot; d.frame = s.frame; d.mesh = s.mesh;
}

static int pair_class(int sa, int sb) {
    if (sa == TMK_MESH || sb == TMK_MESH) return PC_MESH;
    if (sa == TMK_SPHERE || sb == TMK_SPHERE) return PC_CLOSED;
    if (sa == TMK_BOX && sb == TMK_BOX) return PC_SAT;
    return PC_GJK;
}

static void ensure_image(struct aa_rx_cl *cl, size_t n_sub, const aa_rx_config_id *sub_ids, size_t n_q,
                         const double *q_base) {
    const struct aa_rx_sg *sg = cl->sg;
    ImageDev &im = cl->img;
    {
        int cur = -1;
        CK(cudaGetDevice(&cur));
        if (cur != cl->dev) tmk_die("batch: the collision context lives on device %d but device %d is current", cl->dev, cur);
    }
    if (n_q != cl->n_q) tmk_die("batch: n_q = %zu, scene graph has %zu", n_q, cl->n_q);
    if (n_sub < 1 || n_sub > TMK_MAXSUB) tmk_die("batch: n_sub = %zu outside 1..%d", n_sub, TMK_MAXSUB);
    bool same = im.valid && im.allow_serial == cl->allow_serial && im.sub.size() == n_sub &&
                0 == memcmp(im.sub.data(), sub_ids, sizeof(int) * n_sub) &&
                0 == memcmp(im.q_base.data(), q_base, sizeof(double) * n_q);
    if (same) return;
    HostSpan hs("ensure_image");
    for (size_t k = 0; k < n_sub; k++)
        if (sub_ids[k] < 0 || (size_t)sub_ids[k] >= n_q) tmk_die("batch: bad configuration id %d", sub_ids[k]);
    im.sub.assign(sub_ids, sub_ids + n_sub);
    im.q_base.assign(q_base, q_base + n_q);
    im.allow_serial = cl->allow_serial;

    size_t n_f = cl->n_f;
    std::vector<double> rel(7 * n_f), ab(7 * n_f);
    tmk_dev_fk64(sg, q_base, rel.data(), ab.data()); /* static poses come from the GPU FK */
    hs.mark("fk64");

    std::vector<int> slot(n_f, -1), is_joint(n_f, 0);
    std::vector<double> fold(7 * n_f, 0.0);
    std::vector<DJoint<double>> joints;
    for (size_t i = 0; i < n_f; i++) {
        int p = sg->parent[i];
        int qs = -1;
        if (sg->type[i] != TMK_FIXED)
            for (size_t k = 0; k < n_sub; k++)
                if (sub_ids[k] == sg->qidx[i]) qs = (int)k;
        if (qs >= 0) {
            DJoint<double> J;
            memset(&J, 0, sizeof(J));
            const double *E = sg->E + 7 * i;
            if (p < 0) { memcpy(J.E, E, sizeof(double) * 7); J.parent = -1; }
            else if (slot[p] < 0) { tmk_tfmul(&ab[7 * p], E, J.E); J.parent = -1; }
            else { tmk_tfmul(&fold[7 * p], E, J.E); J.parent = slot[p]; }
            memcpy(J.ax, sg->axis + 3 * i, sizeof(double) * 3);
            J.off = sg->offset[i];
            J.type = sg->type[i];
            J.qslot = qs;
            slot[i] = (int)joints.size();
            is_joint[i] = 1;
            joints.push_back(J);
            double id7[7] = {0, 0, 0, 1, 0, 0, 0};
            memcpy(&fold[7 * i], id7, sizeof(id7));
        } else if (p >= 0 && slot[p] >= 0) {
            slot[i] = slot[p];
            tmk_tfmul(&fold[7 * p], &rel[7 * i], &fold[7 * i]);
        }
    }
    if (joints.empty()) tmk_die("batch: none of the sub-configuration variables moves a frame");
    if (joints.size() > TMK_MAXJ) tmk_die("batch: %zu moving joints > %d", joints.size(), TMK_MAXJ);

    /* joints whose children do not all follow them directly keep their pose in the joint slab */
    int n_jslab = 0;
    for (size_t k = 0; k < joints.size(); k++)
        if (joints[k].parent >= 0 && joints[k].parent != (int)k - 1 && joints[joints[k].parent].pad == 0)
            joints[joints[k].parent].pad = ++n_jslab;

    /* moving geometry sorted by joint slot (the kernel walks it joint by joint), static geometry as found */
    std::vector<DGeom<double>> mg, sgv;
    std::vector<int> g2m(cl->n_g, -1), g2s(cl->n_g, -1);
    std::vector<int> jg_begin(joints.size() + 1, 0);
    for (size_t j = 0; j < joints.size(); j++) {
        jg_begin[j] = (int)mg.size();
        for (size_t g = 0; g < cl->n_g; g++) {
            int f = cl->geoms[g].frame;
            if (slot[f] != (int)j) continue;
            DGeom<double> G = cl->geoms[g];
            double L[7];
            memcpy(L, G.tf, sizeof(L));
            tmk_tfmul(&fold[7 * f], L, G.tf);
            G.slot = slot[f];
            g2m[g] = (int)mg.size();
            mg.push_back(G);
        }
    }
    jg_begin[joints.size()] = (int)mg.size();
    for (size_t g = 0; g < cl->n_g; g++) {
        int f = cl->geoms[g].frame;
        if (slot[f] >= 0) continue;
        DGeom<double> G = cl->geoms[g];
        double L[7];
        memcpy(L, G.tf, sizeof(L));
        tmk_tfmul(&ab[7 * f], L, G.tf);
        G.slot = -1;
        g2s[g] = (int)sgv.size();
        sgv.push_back(G);
    }
    if (mg.size() > TMK_MAXMG) tmk_die("batch: %zu moving geometries > %d", mg.size(), TMK_MAXMG);
    if (sgv.size() > 65535) tmk_die("batch: too many static geometries");

    /* candidate pairs: every (moving, other) pair that is not same-frame / allowed.  A pair of two
     * moving geometries is owned by the later one (its partner's pose already exists when the
     * kernel reaches it).  Grouped by owner, then by narrowphase class. */
    /* reach of every moving geometry: it stays inside the ball (centre, reach) whatever the varied
     * joints do (centre = origin of the chain's first varied joint; every further joint origin adds its
     * offset, prismatic joints their travel).  Static partners outside that ball can never be touched and
     * are dropped from the list - verdicts are unaffected. */
    std::vector<double> j_cx(joints.size() * 3, 0.0), j_reach(joints.size(), 0.0);
    for (size_t k = 0; k < joints.size(); k++) {
        const DJoint<double> &J = joints[k];
        double travel = 0;
        if (J.type == TMK_PRISMATIC) {
            int cid = sub_ids[J.qslot];
            double lo = sg->lim_min[cid], hi = sg->lim_max[cid];
            double an = sqrt(J.ax[0] * J.ax[0] + J.ax[1] * J.ax[1] + J.ax[2] * J.ax[2]);
            travel = (std::isnan(lo) || std::isnan(hi)) ? INFINITY : an * std::max(fabs(lo + J.off), fabs(hi + J.off));
        }
        if (J.parent < 0) {
            for (int c = 0; c < 3; c++) j_cx[3 * k + c] = J.E[4 + c];
            j_reach[k] = travel;
        } else {
            for (int c = 0; c < 3; c++) j_cx[3 * k + c] = j_cx[3 * J.parent + c];
            j_reach[k] = j_reach[J.parent] + sqrt(J.E[4] * J.E[4] + J.E[5] * J.E[5] + J.E[6] * J.E[6]) + travel;
        }
    }
    auto unreachable = [&](const DGeom<double> &A, const DGeom<double> &B) {
        const double *c = &j_cx[3 * A.slot];
        double reach = j_reach[A.slot] + sqrt(A.tf[4] * A.tf[4] + A.tf[5] * A.tf[5] + A.tf[6] * A.tf[6]) + A.brad;
        if (A.shape == TMK_MESH) reach += sqrt(A.dim[0] * A.dim[0] + A.dim[1] * A.dim[1] + A.dim[2] * A.dim[2]);
        if (!(reach < 1e9)) return false;
        double d[3] = {c[0] - B.tf[4], c[1] - B.tf[5], c[2] - B.tf[6]}, dist;
        if (B.shape == TMK_BOX) {
            double qi[4] = {-B.tf[0], -B.tf[1], -B.tf[2], B.tf[3]}, l[3];
            tmk_qrot(qi, d, l);
            double e2 = 0;
            for (int k = 0; k < 3; k++) { double e = fabs(l[k]) - B.dim[k]; if (e > 0) e2 += e * e; }
            dist = sqrt(e2);
        } else {
            double off = 0;
            if (B.shape == TMK_MESH) off = sqrt(B.dim[0] * B.dim[0] + B.dim[1] * B.dim[1] + B.dim[2] * B.dim[2]);
            dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) - B.brad - off;
        }
        return dist > reach * (1 + 1e-9) + 1e-4;
    };

    struct P { uint32_t w; int fa, fb, owner, seg, aligned; };
    std::vector<P> pairs;
    for (size_t a = 0; a < cl->n_g; a++) {
        if (g2m[a] < 0) continue;
        for (size_t b = 0; b < cl->n_g; b++) {
            if (b == a) continue;
            if (g2m[b] >= 0 && g2m[b] > g2m[a]) continue; /* owned by b */
            int fa = cl->geoms[a].frame, fb = cl->geoms[b].frame;
            if (!pair_live(cl, fa, fb)) continue;
            if (g2m[b] < 0 && unreachable(mg[g2m[a]], sgv[g2s[b]])) continue;
            int cls = pair_class(cl->geoms[a].shape, cl->geoms[b].shape);
            P p;
            p.w = g2m[b] >= 0 ? pair_pack(g2m[a], g2m[b], 0, cls, cl->geoms[b].shape)
                              : pair_pack(g2m[a], g2s[b], 1, cls, cl->geoms[b].shape);
            p.fa = fa; p.fb = fb; p.owner = g2m[a];
            /* segment 1 (exact sphere-to-oriented-box distance, 2x the cost of a sphere test) pays off for
             * large flat boxes and mesh boxes; a small box is as well served by its bounding sphere */
            const bool big_box = (cl->geoms[b].shape == TMK_BOX && cl->geoms[b].brad > 0.2) || cl->geoms[b].shape == TMK_MESH;
            p.seg = g2m[b] >= 0 ? 2 : (big_box ? 1 : 0);
            /* a box segment entry whose box is aligned with the world axes (tables, walls): the wavefront broadphase
             * tests those as centre + half extents; they come first in the segment */
            p.aligned = 0;
            if (p.seg == 1) {
                const double *t = sgv[g2s[b]].tf;
                const double x = t[0], y = t[1], z = t[2], ww = t[3];
                const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww),
                                     2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww),
                                     2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)};
                p.aligned = 1;
                for (double r : R)
                    if (!(fabs(r) < 1e-9 || fabs(r) > 1 - 1e-9)) p.aligned = 0;
            }
            pairs.push_back(p);
        }
    }
    std::stable_sort(pairs.begin(), pairs.end(), [](const P &x, const P &y) {
        if (x.owner != y.owner) return x.owner < y.owner;
        if (x.seg != y.seg) return x.seg < y.seg;
        if (x.aligned != y.aligned) return x.aligned > y.aligned;
        return (x.w >> 28) < (y.w >> 28);
    });
    im.h_pairs.clear(); im.pair_fa.clear(); im.pair_fb.clear();
    for (auto &p : pairs) { im.h_pairs.push_back(p.w); im.pair_fa.push_back(p.fa); im.pair_fb.push_back(p.fb); }
    if (im.h_pairs.size() >= (1u << 27)) tmk_die("batch: too many candidate pairs");

    /* large scenes: small static geometry goes into a uniform grid and leaves the pipeline's pair list */
    const double GRID_RMAX = 0.25;
    const double GRID_H = getenv("TMK_GRID_H") ? atof(getenv("TMK_GRID_H")) : 0.25; /* cell size (0.2-0.3 measured best on 512 obstacles); env = tuning aid */
    std::vector<uint8_t> gridded(sgv.size(), 0);
    size_t n_gridded = 0;
    for (size_t b = 0; b < sgv.size(); b++)
        if (sgv[b].shape != TMK_MESH && sgv[b].brad <= GRID_RMAX) { gridded[b] = 1; n_gridded++; }
    const bool grid_on = n_gridded >= 96 && !getenv("TMK_NO_GRID");
    if (!grid_on) std::fill(gridded.begin(), gridded.end(), 0);
    StaticGrid grid;
    memset(&grid, 0, sizeof(grid));
    std::vector<int> cell_start(2, 0);
    std::vector<unsigned short> cell_items(1, 0);
    const int wps = (int)(sgv.size() + 31) / 32;
    std::vector<uint32_t> live_ms((size_t)std::max<size_t>(1, mg.size() * (size_t)wps), 0u);
    if (grid_on) {
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (size_t b = 0; b < sgv.size(); b++)
            if (gridded[b])
                for (int c = 0; c < 3; c++) { lo[c] = std::min(lo[c], sgv[b].tf[4 + c]); hi[c] = std::max(hi[c], sgv[b].tf[4 + c]); }
        int dims[3];
        for (int c = 0; c < 3; c++) dims[c] = std::max(1, std::min(64, (int)floor((hi[c] - lo[c]) / GRID_H) + 1));
        grid.on = 1;
        grid.nx = dims[0]; grid.ny = dims[1]; grid.nz = dims[2];
        grid.ox = (float)lo[0]; grid.oy = (float)lo[1]; grid.oz = (float)lo[2];
        grid.inv_h = (float)(1.0 / GRID_H);
        grid.r_max = (float)(GRID_RMAX * (1 + 1e-6) + 1e-5); /* + slack for the float cell arithmetic */
        grid.wps = wps;
        auto cell_of = [&](const DGeom<double> &G) {
            /* the same float arithmetic as the kernel's query (clamped), so that a geometry is always inside
             * the cell range computed for a query sphere that reaches its centre */
            int ix[3];
            const float o[3] = {grid.ox, grid.oy, grid.oz};
            for (int c = 0; c < 3; c++) {
                int v = (int)floorf(((float)G.tf[4 + c] - o[c]) * grid.inv_h);
                ix[c] = std::max(0, std::min(dims[c] - 1, v));
            }
            return (ix[2] * dims[1] + ix[1]) * dims[0] + ix[0];
        };
        const size_t n_cell = (size_t)dims[0] * dims[1] * dims[2];
        cell_start.assign(n_cell + 1, 0);
        for (size_t b = 0; b < sgv.size(); b++)
            if (gridded[b]) cell_start[cell_of(sgv[b]) + 1]++;
        for (size_t k = 0; k < n_cell; k++) cell_start[k + 1] += cell_start[k];
        cell_items.assign(n_gridded + 1, 0);
        std::vector<int> fill(cell_start.begin(), cell_start.end() - 1);
        for (size_t b = 0; b < sgv.size(); b++)
            if (gridded[b]) cell_items[fill[cell_of(sgv[b])]++] = (unsigned short)b;
        for (auto &p : pairs)
            if (((p.w >> 24) & 1) && gridded[(p.w >> 8) & 0xffff])
                live_ms[(size_t)p.owner * wps + (((p.w >> 8) & 0xffff) >> 5)] |= 1u << (((p.w >> 8) & 0xffff) & 31);
    }
    /* reach lists for the wavefront broadphase (kernels.cuh StaticGrid::r_*) */
    std::vector<int> r_start(2, 0);
    std::vector<uint2> r_c(1, make_uint2(0, 0));
    std::vector<unsigned short> r_id(1, 0);
    if (grid_on) {
        double ra_max = 0;
        for (auto &G : mg) ra_max = std::max(ra_max, G.brad * (1 + 1e-6));
        ra_max += 1e-4; /* the kernel's pad (4 eps) and its float cell arithmetic */
        double H = getenv("TMK_REACH_H") ? atof(getenv("TMK_REACH_H")) : 0.125; /* tuning aid */
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        std::vector<double> rb(sgv.size(), 0.0);
        for (size_t b = 0; b < sgv.size(); b++) {
            if (!gridded[b]) continue;
            rb[b] = (double)(float)(sgv[b].brad * (1 + 1e-6)) * (1 + 1e-6); /* >= the float radius the quick tables hold */
            for (int c = 0; c < 3; c++) {
                lo[c] = std::min(lo[c], sgv[b].tf[4 + c] - ra_max - rb[b] - 1e-3);
                hi[c] = std::max(hi[c], sgv[b].tf[4 + c] + ra_max + rb[b] + 1e-3);
            }
        }
        int dims[3];
        for (;;) {
            size_t cells = 1;
            for (int c = 0; c < 3; c++) { dims[c] = std::max(1, (int)ceil((hi[c] - lo[c]) / H)); cells *= (size_t)dims[c]; }
            if (cells <= (size_t(1) << 20)) break;
            H *= 1.26; /* halve the cell count */
        }
        const size_t n_cell = (size_t)dims[0] * dims[1] * dims[2];
        /* quantisation: coordinates relative to the cell centre are below R + h sqrt(3)/2 in magnitude */
        double vmax = 0;
        for (size_t b = 0; b < sgv.size(); b++)
            if (gridded[b]) vmax = std::max(vmax, ra_max + rb[b] + H);
        const double ulp = ldexp(1.0, (int)floor(log2(std::max(vmax, 1e-3))) - 10); /* half precision: 10 mantissa bits */
        const double q_err = 0.5 * ulp * sqrt(3.0) + 1e-6;
        auto half_up = [](double v) { /* smallest half >= v (v > 0) */
            __half h = __float2half_rn((float)v);
            unsigned short u = __half_as_ushort(h);
            if ((double)__half2float(h) < v) u++;
            return u;
        };
        auto half_rn = [](double v) { return (unsigned short)__half_as_ushort(__float2half_rn((float)v)); };
        for (int pass = 0; pass < 2; pass++) {
            std::vector<int> fill;
            if (pass == 0) r_start.assign(n_cell + 1, 0);
            else {
                for (size_t k = 0; k < n_cell; k++) r_start[k + 1] += r_start[k];
                r_c.assign((size_t)r_start[n_cell] + 1, make_uint2(0, 0));
                r_id.assign((size_t)r_start[n_cell] + 1, 0);
                fill.assign(r_start.begin(), r_start.end() - 1);
            }
            for (size_t b = 0; b < sgv.size(); b++) {
                if (!gridded[b]) continue;
                const double R = ra_max + rb[b] + 1e-4, *c = &sgv[b].tf[4];
                int i0[3], i1[3];
                for (int a = 0; a < 3; a++) {
                    i0[a] = std::max(0, (int)floor((c[a] - R - lo[a]) / H));
                    i1[a] = std::min(dims[a] - 1, (int)floor((c[a] + R - lo[a]) / H));
                }
                for (int z = i0[2]; z <= i1[2]; z++)
                    for (int y = i0[1]; y <= i1[1]; y++)
                        for (int x = i0[0]; x <= i1[0]; x++) {
                            const int ix[3] = {x, y, z};
                            double d2 = 0, rel[3];
                            for (int a = 0; a < 3; a++) {
                                const double bl = lo[a] + ix[a] * H, bh = bl + H;
                                const double e = std::max(std::max(bl - c[a], c[a] - bh), 0.0);
                                d2 += e * e;
                                rel[a] = c[a] - (bl + 0.5 * H);
                            }
                            if (d2 > R * R) continue;
                            const size_t cell = ((size_t)z * dims[1] + y) * dims[0] + x;
                            if (pass == 0) { r_start[cell + 1]++; continue; }
                            const int k = fill[cell]++;
                            r_c[k] = make_uint2((uint32_t)half_rn(rel[0]) | ((uint32_t)half_rn(rel[1]) << 16),
                                                (uint32_t)half_rn(rel[2]) | ((uint32_t)half_up(rb[b] + q_err) << 16));
                            r_id[k] = (unsigned short)b;
                        }
            }
        }
        grid.r_on = 1;
        grid.r_nx = dims[0]; grid.r_ny = dims[1]; grid.r_nz = dims[2];
        grid.r_ox = (float)lo[0]; grid.r_oy = (float)lo[1]; grid.r_oz = (float)lo[2];
        grid.r_h = (float)H; grid.r_inv_h = (float)(1.0 / H);
        if (getenv("TMK_DEBUG_IMAGE"))
            fprintf(stderr, "tmkcl reach lists: %d x %d x %d cells of %.3f m, %d entries (%.1f per cell), query radius <= %.3f m\n", dims[0], dims[1],
                    dims[2], H, r_start[n_cell], (double)r_start[n_cell] / (double)n_cell, ra_max);
    }
    std::vector<P> tpairs;
    for (auto &p : pairs)
        if (!(((p.w >> 24) & 1) && gridded[(p.w >> 8) & 0xffff])) tpairs.push_back(p);
    std::vector<uint32_t> h_tpairs;
    for (auto &p : tpairs) h_tpairs.push_back(p.w);
    std::vector<int> pa_begin(3 * mg.size() + 1, 0);
    if (getenv("TMK_DEBUG_IMAGE")) {
        size_t seg_n[3] = {0, 0, 0};
        for (auto &p : tpairs) seg_n[p.seg]++;
        fprintf(stderr, "tmkcl image: %zu joints, %zu moving / %zu static geometries, %zu pairs (%zu in the pipeline list: %zu sphere, %zu box, %zu moving), grid %s (%zu obstacles)\n",
                joints.size(), mg.size(), sgv.size(), pairs.size(), tpairs.size(), seg_n[0], seg_n[1], seg_n[2], grid_on ? "on" : "off", grid_on ? n_gridded : (size_t)0);
    }
    for (auto &p : tpairs) pa_begin[3 * p.owner + p.seg + 1]++;
    for (size_t k = 0; k < 3 * mg.size(); k++) pa_begin[k + 1] += pa_begin[k];

    /* world-frame quick tables of the static geometry */
    std::vector<SQuick> sq(sgv.size());
    for (size_t k = 0; k < sgv.size(); k++) {
        const DGeom<double> &G = sgv[k];
        SQuick &Q = sq[k];
        double x = G.tf[0], y = G.tf[1], z = G.tf[2], w = G.tf[3];
        double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                       2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        Q.a0x = (float)R[0]; Q.a0y = (float)R[3]; Q.a0z = (float)R[6];
        Q.a1x = (float)R[1]; Q.a1y = (float)R[4]; Q.a1z = (float)R[7];
        Q.a2x = (float)R[2]; Q.a2y = (float)R[5]; Q.a2z = (float)R[8];
        double c[3] = {G.tf[4], G.tf[5], G.tf[6]};
        if (G.shape == TMK_MESH)
            for (int r = 0; r < 3; r++) c[r] += R[3 * r] * G.dim[0] + R[3 * r + 1] * G.dim[1] + R[3 * r + 2] * G.dim[2];
        Q.cx = (float)c[0]; Q.cy = (float)c[1]; Q.cz = (float)c[2];
        Q.brad = (float)(G.brad * (1 + 1e-6));
        Q.h0 = (float)G.dim[0]; Q.h1 = (float)G.dim[1]; Q.h2 = (float)G.dim[2];
        if (G.shape == TMK_MESH) { /* broadphase sees a static mesh as the oriented box of its root node */
            const BvhNode &root = cl->mesh_root[G.mesh];
            Q.h0 = 0.5f * (root.hi[0] - root.lo[0]) * (1 + 1e-6f) + 1e-6f;
            Q.h1 = 0.5f * (root.hi[1] - root.lo[1]) * (1 + 1e-6f) + 1e-6f;
            Q.h2 = 0.5f * (root.hi[2] - root.lo[2]) * (1 + 1e-6f) + 1e-6f;
        }
    }

    /* wavefront broadphase tables in t_pairs order: ready queue keys, world-aligned boxes, end of the aligned part */
    std::vector<uint32_t> t_keys(tpairs.size() + 1, 0u);
    std::vector<float4> ab4(2 * tpairs.size() + 2, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<int> pa_aabb(mg.size() + 1, 0);
    for (size_t a = 0; a < mg.size(); a++) pa_aabb[a] = pa_begin[3 * a + 1];
    for (size_t k = 0; k < tpairs.size(); k++) {
        const uint32_t w = tpairs[k].w;
        const int b = (int)((w >> 8) & 0xffff), st = (int)((w >> 24) & 1);
        t_keys[k] = ((uint32_t)tpairs[k].owner << 17) | ((uint32_t)st << 16) | (uint32_t)b; /* ITEM_KEY(item_pack(0, a, b, st)) */
        if (tpairs[k].seg == 1 && tpairs[k].aligned) {
            const SQuick &Q = sq[b];
            const float R[9] = {Q.a0x, Q.a1x, Q.a2x, Q.a0y, Q.a1y, Q.a2y, Q.a0z, Q.a1z, Q.a2z}; /* row i: world component i of the axes */
            float h[3];
            for (int i = 0; i < 3; i++) {
                const double e = fabs((double)R[3 * i]) * Q.h0 + fabs((double)R[3 * i + 1]) * Q.h1 + fabs((double)R[3 * i + 2]) * Q.h2;
                h[i] = nextafterf((float)(e * (1 + 1e-6)), INFINITY); /* outward: the test stays conservative */
            }
            ab4[2 * k] = make_float4(Q.cx, Q.cy, Q.cz, 0.f);
            ab4[2 * k + 1] = make_float4(h[0], h[1], h[2], 0.f);
            pa_aabb[tpairs[k].owner] = (int)k + 1;
        }
    }
    hs.mark("host lists");
    /* hoisted static-static pairs, evaluated once (double) */
    std::vector<uint2> sp;
    for (size_t a = 0; a < cl->n_g; a++)
        for (size_t b = a + 1; b < cl->n_g; b++)
            if (g2m[a] < 0 && g2m[b] < 0 && pair_live(cl, cl->geoms[a].frame, cl->geoms[b].frame))
                sp.push_back(make_uint2((unsigned)a, (unsigned)b));
    int static_hit = 0;
    if (!sp.empty()) {
        DBuf tmp;
        upload(tmp, sp, cl->stream);
        std::vector<uint8_t> out;
        run_pairs_tf(cl, ab.data(), sp, tmp, out);
        for (uint8_t o : out) static_hit |= o;
        tmp.release();
    }
    hs.mark("static pairs");

    std::vector<DJoint<float>> jf(joints.size());
    for (size_t k = 0; k < joints.size(); k++) cvt_joint(joints[k], jf[k]);
    std::vector<DGeom<float>> mgf(mg.size()), sgf(sgv.size());
    for (size_t k = 0; k < mg.size(); k++) cvt_geom(mg[k], mgf[k]);
    for (size_t k = 0; k < sgv.size(); k++) cvt_geom(sgv[k], sgf[k]);
    upload(im.jd, joints, cl->stream); upload(im.jf, jf, cl->stream);
    upload(im.mgd, mg, cl->stream); upload(im.mgf, mgf, cl->stream);
    upload(im.sgd, sgv, cl->stream); upload(im.sgf, sgf, cl->stream);
    upload(im.pairs, im.h_pairs, cl->stream);
    upload(im.sq, sq, cl->stream);
    upload(im.jg, jg_begin, cl->stream);
    upload(im.pa, pa_begin, cl->stream);
    upload(im.tpairs, h_tpairs, cl->stream);
    upload(im.tkeys, t_keys, cl->stream);
    upload(im.ab4, ab4, cl->stream);
    upload(im.paabb, pa_aabb, cl->stream);
    upload(im.g_cells, cell_start, cl->stream);
    upload(im.g_items, cell_items, cl->stream);
    upload(im.g_live, live_ms, cl->stream);
    {
        std::vector<float4> cell_c4(cell_items.size() + 1);
        for (size_t k = 0; k < cell_items.size(); k++) {
            const SQuick &Q = sq.empty() ? SQuick() : sq[std::min<size_t>(cell_items[k], sq.size() - 1)];
            cell_c4[k] = make_float4(Q.cx, Q.cy, Q.cz, Q.brad);
        }
        cell_c4[cell_items.size()] = make_float4(0.f, 0.f, 0.f, 0.f);
        upload(im.g_c4, cell_c4, cl->stream);
    }
    {
        std::vector<float4> ps4(tpairs.size() + 1);
        /* .w = (bounding radius of the owner + the kernels' pad + bounding radius of the partner)^2, rounded up: the
         * sphere tests of the broadphase compare a squared distance with it (static partners: centre in .xyz) */
        auto reach2 = [](double ra, double rb) {
            const double r = ra + 4.0 * 1.0e-5 + rb, r2 = r * r;
            float f = (float)r2;
            if ((double)f < r2) f = nextafterf(f, INFINITY);
            return f;
        };
        for (size_t k = 0; k < tpairs.size(); k++) {
            uint32_t w = tpairs[k].w;
            const double ra = (double)mgf[tpairs[k].owner].brad;
            if ((w >> 24) & 1) {
                const SQuick &Q = sq[(w >> 8) & 0xffff];
                ps4[k] = make_float4(Q.cx, Q.cy, Q.cz, reach2(ra, (double)Q.brad));
            } else ps4[k] = make_float4(0.f, 0.f, 0.f, reach2(ra, (double)mgf[(w >> 8) & 0xffff].brad));
        }
        upload(im.ps4, ps4, cl->stream);
    }
    CK(cudaStreamSynchronize(cl->stream));
    hs.mark("uploads");

    im.d.joints = (const DJoint<double> *)im.jd.p; im.f.joints = (const DJoint<float> *)im.jf.p;
    im.d.mg = (const DGeom<double> *)im.mgd.p; im.f.mg = (const DGeom<float> *)im.mgf.p;
    im.d.sg = (const DGeom<double> *)im.sgd.p; im.f.sg = (const DGeom<float> *)im.sgf.p;
    im.d.pairs = im.f.pairs = (const uint32_t *)im.pairs.p;
    im.d.meshes = (const MeshRef<double> *)cl->d_meshd.p; im.f.meshes = (const MeshRef<float> *)cl->d_meshf.p;
    im.d.n_joint = im.f.n_joint = (int)joints.size();
    im.d.n_mg = im.f.n_mg = (int)mg.size();
    im.d.n_sg = im.f.n_sg = (int)sgv.size();
    im.d.n_pair = im.f.n_pair = (int)im.h_pairs.size();
    im.d.n_sub = im.f.n_sub = (int)n_sub;
    im.d.static_hit = im.f.static_hit = static_hit;
    im.d.sq = im.f.sq = im.sq.p;
    im.d.jg_begin = im.f.jg_begin = (const int *)im.jg.p;
    im.d.pa_begin = im.f.pa_begin = (const int *)im.pa.p;
    im.d.ps4 = im.f.ps4 = im.ps4.p;
    im.d.t_pairs = im.f.t_pairs = (const uint32_t *)im.tpairs.p;
    im.d.t_keys = im.f.t_keys = (const uint32_t *)im.tkeys.p;
    im.d.ab4 = im.f.ab4 = im.ab4.p;
    im.d.pa_aabb = im.f.pa_aabb = (const int *)im.paabb.p;
    im.d.n_tpair = im.f.n_tpair = (int)h_tpairs.size();
    grid.cell_start = (const int *)im.g_cells.p;
    grid.items = (const unsigned short *)im.g_items.p;
    grid.live = (const uint32_t *)im.g_live.p;
    upload(im.g_rstart, r_start, cl->stream);
    upload(im.g_rc, r_c, cl->stream);
    upload(im.g_rid, r_id, cl->stream);
    CK(cudaStreamSynchronize(cl->stream));
    grid.r_start = (const int *)im.g_rstart.p;
    grid.r_c = (const uint2 *)im.g_rc.p;
    grid.r_id = (const unsigned short *)im.g_rid.p;
    grid.c4 = (const float4 *)im.g_c4.p;
    grid.n_items = grid.on ? (int)n_gridded : 0;
    im.d.grid = im.f.grid = grid;
    {
        TileCfg tmp; /* the flat edge source shares the level source's plan; its kernels need the same attributes */
        if (grid.on) tile_configure(tmp, im.f, n_jslab, k_states_tile<SrcEdgeFlat, SinkEdgeFirst, 256, true>, k_states_tile<SrcEdgeFlat, SinkEdgeFirst, 128, true>);
        else tile_configure(tmp, im.f, n_jslab, k_states_tile<SrcEdgeFlat, SinkEdgeFirst, 256, false>, k_states_tile<SrcEdgeFlat, SinkEdgeFirst, 128, false>);
        im.wcfg_flat = tmp;
    }
    if (grid.on) {
        tile_configure(im.wcfg, im.f, n_jslab, k_states_tile<SrcConfigs, SinkBits, 256, true>,
                       k_states_tile<SrcConfigs, SinkBits, 128, true>);
        tile_configure(im.wcfg_e, im.f, n_jslab, k_states_tile<SrcEdgeLevel, SinkEdgeFirst, 256, true>,
                       k_states_tile<SrcEdgeLevel, SinkEdgeFirst, 128, true>);
    } else {
        tile_configure(im.wcfg, im.f, n_jslab, k_states_tile<SrcConfigs, SinkBits, 256, false>,
                       k_states_tile<SrcConfigs, SinkBits, 128, false>);
        tile_configure(im.wcfg_e, im.f, n_jslab, k_states_tile<SrcEdgeLevel, SinkEdgeFirst, 256, false>,
                       k_states_tile<SrcEdgeLevel, SinkEdgeFirst, 128, false>);
    }
    if (mg.size() > 64) im.wcfg.smem = im.wcfg_e.smem = im.wcfg_flat.smem = -1; /* queue items hold a 6-bit moving-geometry index */
    if (getenv("TMK_DEBUG_IMAGE"))
        fprintf(stderr, "tmkcl tile: configs TILE %d, %d B shared memory, grid %d | edges TILE %d, %d B, grid %d\n", im.wcfg.tile,
                im.wcfg.smem, im.wcfg.grid, im.wcfg_e.tile, im.wcfg_e.smem, im.wcfg_e.grid);
    im.valid = true;
    im.serial++;
    hs.mark("tile_configure");
}

/* ------------------------------------------------------------------ batch launches */
static void stats_begin(struct aa_rx_cl *cl) {
    memset(&cl->stats, 0, sizeof(cl->stats));
    if (cl->stats_on) {
        cl->d_stats.need(sizeof(DevStats));
        CK(cudaMemsetAsync(cl->d_stats.p, 0, sizeof(DevStats), cl->stream));
    }
}

/* the four instantiations of the pipeline kernel: tile size x static grid */
template <class Src, class Sink>
static void launch_tile(const Image<float> &f, const TileCfg &cfg, const Src &src, const Sink &sink, const UncList &u,
                        unsigned long long *n_states, unsigned grid, cudaStream_t s) {
    if (cfg.tile == 256) {
        if (f.grid.on) k_states_tile<Src, Sink, 256, true><<<grid, 256, cfg.smem, s>>>(f, cfg, src, sink, u, n_states);
        else k_states_tile<Src, Sink, 256, false><<<grid, 256, cfg.smem, s>>>(f, cfg, src, sink, u, n_states);
    } else {
        if (f.grid.on) k_states_tile<Src, Sink, 128, true><<<grid, 128, cfg.smem, s>>>(f, cfg, src, sink, u, n_states);
        else k_states_tile<Src, Sink, 128, false><<<grid, 128, cfg.smem, s>>>(f, cfg, src, sink, u, n_states);
    }
}

/* the deferred narrowphase: convex pairs, mesh traversal, open triangles (in this order: the triangle
 * list is filled by the traversal) */
template <class Src, class Sink>
static void launch_heavy(const Image<float> &f, const Src &src, const Sink &sink, const UncList &u, unsigned ctas,
                         cudaStream_t s, uint32_t *trav_next, struct aa_rx_cl *cl) {
    /* the convex items (SAT / GJK) are independent of the mesh chain (traversal -> open triangles): they run
     * beside it on the context's second stream and rejoin before the exact re-check.  Inside a stream capture the
     * fork becomes two branches of the graph. */
    static const bool fork = getenv("TMK_NO_FORK") == nullptr; /* A/B aid */
    cudaStream_t sc = fork ? cl->aux_stream : s;
    if (fork) {
        CK(cudaEventRecord(cl->ev_aux[0], s));
        CK(cudaStreamWaitEvent(sc, cl->ev_aux[0], 0));
    }
    k_conv_items<Src, Sink><<<ctas, 128, 0, sc>>>(f, src, sink, u);
    CK(cudaGetLastError());
    if (fork) CK(cudaEventRecord(cl->ev_aux[1], sc));
    /* the tree walks: warp task stacks (default); TMK_TRAVERSE=pool / simple select the earlier kernels (A/B aid) */
    static const int mode = !getenv("TMK_TRAVERSE") ? 0 : !strcmp(getenv("TMK_TRAVERSE"), "pool") ? 1 : !strcmp(getenv("TMK_TRAVERSE"), "simple") ? 2 : 0;
    static const unsigned trav_ctas = getenv("TMK_TRAV_CTAS") ? (unsigned)atoi(getenv("TMK_TRAV_CTAS")) : (mode == 1 ? 7u : 6u);
    if (mode == 2) k_mesh_traverse_simple<Src, Sink><<<ctas, 128, 0, s>>>(f, src, sink, u);
    else if (mode == 1) k_mesh_traverse<Src, Sink><<<148 * trav_ctas, 128, 0, s>>>(f, src, sink, u, trav_next);
    else {
        uint32_t task_cap = TMK_TASK_NQ;
        if (const char *e = getenv("TMK_TASK_CAP")) task_cap = (uint32_t)std::max(32, atoi(e)); /* test aid: the overflow path of k_mesh_tasks */
        k_mesh_tasks<Src, Sink><<<148 * trav_ctas, 128, 0, s>>>(f, src, sink, u, trav_next, task_cap);
    }
    CK(cudaGetLastError());
    k_tri_items<Src, Sink><<<ctas, 128, 0, s>>>(f, src, sink, u);
    CK(cudaGetLastError());
    if (fork) CK(cudaStreamWaitEvent(s, cl->ev_aux[1], 0));
}

static UncList unc_list(struct aa_rx_cl *cl, size_t n_states, cudaStream_t s, bool reset = true) {
    UncList u;
    DBuf &b_unc = cl->d_unc, &b_cnt = cl->d_cnt, &b_mesh = cl->d_mesh, &b_conv = cl->d_conv, &b_tri = cl->d_tri;
    u.cap = (uint32_t)std::min<size_t>(std::max<size_t>(65536, n_states / 2), 0x7fffffffu);
    if (const char *e = getenv("TMK_UNC_CAP")) u.cap = (uint32_t)std::max(1, atoi(e)); /* test aid: forces k_overflow_all */
    b_unc.need(sizeof(uint4) * (size_t)u.cap);
    b_cnt.need(512);
    u.items = (uint4 *)b_unc.p;
    u.n = (uint32_t *)b_cnt.p;
    u.overflow = (uint32_t *)b_cnt.p + 1;
    /* per-CTA spill area of the broadphase queue (only touched when a shared-memory region fills up) */
    u.spill_cap = 1u << 15;
    if (const char *e = getenv("TMK_SPILL_CAP")) u.spill_cap = (uint32_t)std::max(0, atoi(e)); /* test aid */
    const size_t ctas = (size_t)std::max(std::max(cl->img.wcfg.grid, cl->img.wcfg_e.grid), 1);
    cl->d_spill.need(sizeof(uint32_t) * ctas * (size_t)std::max<uint32_t>(u.spill_cap, 1));
    u.spill = (uint32_t *)cl->d_spill.p;
    /* exported mesh items: room for 1.5 per state (more than that: the exact re-check takes them) */
    u.mesh_cap = (uint32_t)std::min<size_t>(std::max<size_t>(65536, n_states + n_states / 2), 0x3fffffffu);
    if (const char *e = getenv("TMK_MESH_CAP")) u.mesh_cap = (uint32_t)std::max(0, atoi(e)); /* test aid */
    b_mesh.need(sizeof(float4) * TMK_MESH_REC * (size_t)std::max<uint32_t>(u.mesh_cap, 1));
    u.mesh_items = (float4 *)b_mesh.p;
    /* d_cnt words: [0] undecided, [1] overflow, [2..3] states, [4] mesh items, [5] convex items, [6] triangles, [7] work counter of k_mesh_traverse, [26..27] tile counter of k_states_tile,
     * [8+L] alive edges of level L, [32+L] / [48+L] / [64+L] mesh / convex / triangle items of edge level L */
    u.mesh_n = (uint32_t *)b_cnt.p + 4;
    u.conv_cap = (uint32_t)std::min<size_t>(std::max<size_t>(65536, n_states), 0x3fffffffu);
    u.tri_cap = u.conv_cap;
    if (const char *e = getenv("TMK_MESH_CAP")) u.conv_cap = u.tri_cap = u.mesh_cap;
    b_conv.need(sizeof(float4) * TMK_MESH_REC * (size_t)std::max<uint32_t>(u.conv_cap, 1));
    b_tri.need(sizeof(uint2) * (size_t)std::max<uint32_t>(u.tri_cap, 1));
    u.conv_items = (float4 *)b_conv.p;
    u.conv_n = (uint32_t *)b_cnt.p + 5;
    u.tri_items = (uint2 *)b_tri.p;
    u.tri_n = (uint32_t *)b_cnt.p + 6;
    u.tile_ctr = (uint32_t *)b_cnt.p + 26; /* words 26, 27 */
    if (reset) CK(cudaMemsetAsync(b_cnt.p, 0, 512, s)); /* after every allocation: capturable */
    return u;
}

/* ---- configuration batches: the wavefront pipeline (wave.cuh) instead of the tile kernel ---- */
static size_t wave_chunk() {
    static const size_t c = getenv("TMK_WAVE_CHUNK") ? (size_t)std::max(4096L, atol(getenv("TMK_WAVE_CHUNK"))) & ~(size_t)31 : (size_t(1) << 20); /* tuning aid */
    return c;
}
#define TMK_WAVE_CHUNK wave_chunk()
static bool wave_wanted(const struct aa_rx_cl *cl, size_t n_states, bool host_chunk = false) {
    /* behind the chunk-wise copies of the host-buffer call the tile kernel's single launch per chunk wins on small
     * scenes (measured 1.47 vs 1.54 ms per 2^20 double rows); large scenes stay with the wavefront kernels, which
     * are several times faster there, on larger chunks */
    if (host_chunk && !cl->img.f.grid.on) return false;
    static const bool off = getenv("TMK_NO_WAVE") != nullptr; /* A/B aid and test hook */
    /* every batch large enough to fill the GPU per kernel; small batches (a planner's frontier) and the swept-edge
     * levels stay on the tile kernel, which needs one launch and no scratch in HBM */
    return n_states >= 4096 && cl->img.f.n_mg <= 64 && !off && !getenv("TMK_NO_WAVE_NOW");
}
static WaveBuf wave_buffers(struct aa_rx_cl *cl, size_t n_states) {
    WaveBuf w;
    const size_t cap = std::min(n_states, TMK_WAVE_CHUNK);
    const size_t capr = (cap + 31) & ~(size_t)31;
    w.cap = (uint32_t)capr;
    cl->d_wpose.need(sizeof(float) * 7 * (size_t)std::max(1, cl->img.f.n_mg) * capr);
    cl->d_whit.need(capr);
    w.queue_cap = (uint32_t)std::min<size_t>(capr * 64, 0x7ffffff0u);
    w.pend_cap = (uint32_t)std::min<size_t>(capr * 8, 0x7ffffff0u);
    if (const char *e = getenv("TMK_WAVE_QCAP")) w.queue_cap = w.pend_cap = (uint32_t)std::max(1, atoi(e)); /* test aid */
    cl->d_wqueue.need(sizeof(uint2) * (size_t)w.queue_cap);
    cl->d_wpend.need(sizeof(uint2) * (size_t)w.pend_cap);
    w.pose = (float *)cl->d_wpose.p;
    w.hit = (uint8_t *)cl->d_whit.p;
    w.queue = (uint2 *)cl->d_wqueue.p;
    w.pend = (uint2 *)cl->d_wpend.p;
    w.qcnt = w.pcnt = nullptr; /* set by the caller: words 16 .. 25 of the batch's counter block */
    return w;
}
/* states [begin, end) of the batch at dQ through the wavefront kernels; exported items join the batch's lists */
static size_t launch_wave(struct aa_rx_cl *cl, const void *dQ, int layout, size_t ld, size_t begin, size_t end,
                          uint32_t *d_bits, const UncList &u, cudaStream_t s) {
    ImageDev &im = cl->img;
    WaveBuf w = wave_buffers(cl, end - begin);
    w.pcnt = u.n + 16;
    const size_t smem = sizeof(uint2) * 8 * TMK_WAVE_STAGE;
    size_t launches = 0;
    /* the broadphase and the quick certificates in stages over the moving geometries, deepest first: a state a stage
     * condemns is skipped by the later ones (both kernels test the state's flag first) */
    const int n_mg = im.f.n_mg;
    const int want = getenv("TMK_WAVE_STAGES") ? atoi(getenv("TMK_WAVE_STAGES")) : 0; /* tuning / test aid (read per call) */
    const int n_stage = std::max(1, std::min(std::min(8, n_mg), want > 0 ? want : (n_mg >= 16 ? 3 : n_mg >= 6 ? 2 : 1)));
    static const bool shallow = getenv("TMK_WAVE_ORDER") && !strcmp(getenv("TMK_WAVE_ORDER"), "shallow"); /* A/B aid */
    for (size_t c0 = begin; c0 < end; c0 += TMK_WAVE_CHUNK) {
        const size_t c1 = std::min(end, c0 + TMK_WAVE_CHUNK);
        const uint32_t n = (uint32_t)(c1 - c0);
        SrcConfigs src;
        src.p = dQ; src.layout = layout; src.ld = ld; src.n = n; src.s_begin = c0;
        CK(cudaMemsetAsync(u.n + 16, 0, 40, s)); /* pend fill and the queue fills of up to eight stages */
        const unsigned gx = (n + 255) / 256;
        w.qcnt = u.n + 18;
        k_wave_fk<SrcConfigs><<<gx, 256, smem, s>>>(im.f, src, w, u);
        launches++;
        for (int g = 0; g < n_stage; g++) {
            /* index ranges, the highest first: geometries are listed in chain order, the deep ones sweep the most space */
            const int gg = shallow ? g : n_stage - 1 - g;
            const int a0 = (int)((long)n_mg * gg / n_stage), a1 = (int)((long)n_mg * (gg + 1) / n_stage);
            w.qcnt = u.n + 18 + g; /* the moving-moving pairs k_wave_fk voted on ride with stage 0 */
            /* two states per thread: measured 0.674 -> 0.633 ms per 2^20 (three or four: 64 registers, slower) */
            static const int broad_k = getenv("TMK_BROAD_K") ? atoi(getenv("TMK_BROAD_K")) : 2; /* A/B aid */
            if (broad_k == 2) k_wave_broad<2><<<dim3((n + 511) / 512, (unsigned)(a1 - a0)), 256, smem, s>>>(im.f, w, n, u, c0, a0);
            else k_wave_broad<1><<<dim3(gx, (unsigned)(a1 - a0)), 256, smem, s>>>(im.f, w, n, u, c0, a0);
            k_wave_quick<<<148 * 8, 256, 0, s>>>(im.f, w, u, c0);
            launches += 2;
        }
        k_wave_export<<<148 * 16, 256, 0, s>>>(im.f, w, u, c0);
        k_wave_bits<<<gx, 256, 0, s>>>(w, n, d_bits + c0 / 32);
        CK(cudaGetLastError());
        launches += 2;
    }
    return launches;
}

/* the caller may be capturing the stream into a graph of its own: then no nested capture / replay */
static bool stream_is_capturing(cudaStream_t s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
    return st != cudaStreamCaptureStatusNone;
}

/* dQ / d_bits address the whole batch of n_cfg configurations; [chunk_begin, chunk_end) is the part this
 * call evaluates (chunk_end == 0: all of it).  Chunks must be issued in order on the same stream. */
static void launch_configs(struct aa_rx_cl *cl, size_t n_cfg, const void *dQ, int layout, size_t ld,
                           uint32_t *d_bits, cudaStream_t s, size_t chunk_begin = 0, size_t chunk_end = 0) {
    ImageDev &im = cl->img;
    if (n_cfg > 0xfffffff0u) tmk_die("batch: more than 2^32 configurations in one call");
    uint8_t *pairhit = nullptr;
    if (cl->track) {
        cl->d_pairhit.need(im.h_pairs.size() + 16);
        CK(cudaMemsetAsync(cl->d_pairhit.p, 0, im.h_pairs.size() + 16, s));
        pairhit = (uint8_t *)cl->d_pairhit.p;
    }
    DevStats *st = cl->stats_on ? (DevStats *)cl->d_stats.p : nullptr;
    if (cl->use_tile && im.wcfg.smem > 0 && !pairhit && !st) {
        /* fast path: FP32 tile pipeline, then the exact re-check of the undecided pairs.  chunk_* let the
         * host-buffer entry point run the pipeline chunk by chunk behind the copies (one undecided list
         * and one re-check for the whole batch) */
        /* a whole (unchunked) batch with arguments seen before is replayed from a CUDA graph, like the swept
         * edges: one driver call per batch keeps a ~1 ms step fed from a busy host */
        const bool whole = chunk_begin == 0 && (chunk_end == 0 || chunk_end == n_cfg);
        bool use_graph = whole && n_cfg > 0 && !cl->timing && !getenv("TMK_NO_GRAPH") && !stream_is_capturing(s);
        uint64_t key[12];
        if (use_graph) {
            UncList u0 = unc_list(cl, n_cfg, s, false); /* allocations happen outside the capture */
            (void)u0;
            if (wave_wanted(cl, n_cfg)) (void)wave_buffers(cl, n_cfg);
            const uint64_t k0[12] = {(uint64_t)n_cfg | (1ull << 62), (uint64_t)(uintptr_t)dQ, 0, (uint64_t)layout * 1315423911u + ld, 0,
                                     (uint64_t)(uintptr_t)d_bits, 0, (uint64_t)(uintptr_t)s, im.serial,
                                     (uint64_t)(uintptr_t)cl->d_unc.p ^ ((uint64_t)(uintptr_t)cl->d_mesh.p << 1) ^ ((uint64_t)(uintptr_t)cl->d_conv.p << 2),
                                     (uint64_t)(uintptr_t)cl->d_wpose.p ^ ((uint64_t)(uintptr_t)cl->d_wqueue.p << 1) ^ ((uint64_t)(uintptr_t)cl->d_wpend.p << 2) ^ ((uint64_t)(uintptr_t)cl->d_whit.p << 3),
                                     (uint64_t)(uintptr_t)cl->d_spill.p ^ ((uint64_t)(uintptr_t)cl->d_tri.p << 1) ^ ((uint64_t)(uintptr_t)cl->d_cnt.p << 2)};
            memcpy(key, k0, sizeof(key));
            for (auto &g : cl->edge_graphs)
                if (0 == memcmp(g.key, key, sizeof(key))) {
                    g.used = ++cl->graph_clock;
                    CK(cudaGraphLaunch(g.exec, s));
                    cl->stats.n_launches += g.n_kernels;
                    return;
                }
            bool seen = false;
            for (auto &k : cl->edge_seen) seen |= 0 == memcmp(k.data(), key, sizeof(key));
            if (!seen) {
                std::array<uint64_t, 12> k;
                memcpy(k.data(), key, sizeof(key));
                if (cl->edge_seen.size() >= 32) cl->edge_seen.erase(cl->edge_seen.begin());
                cl->edge_seen.push_back(k);
                use_graph = false;
            } else CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        }
        if (chunk_begin == 0) cl->cur_unc = unc_list(cl, n_cfg, s);
        UncList u = cl->cur_unc;
        if (n_cfg == 0) return;
        const size_t c_end = chunk_end ? chunk_end : n_cfg;
        SrcConfigs src;
        src.p = dQ; src.layout = layout; src.ld = ld; src.n = c_end - chunk_begin; src.s_begin = chunk_begin;
        SinkBits sink;
        sink.bits = d_bits;
        const int T = im.wcfg.tile;
        unsigned grid = (unsigned)std::min<size_t>((src.n + T - 1) / T, (size_t)im.wcfg.grid);
        cudaEvent_t *ev = chunk_begin == 0 ? (cl->cur_ev = timing_begin(cl, s)) : cl->cur_ev;
        size_t n_main = 1;
        if (wave_wanted(cl, c_end - chunk_begin, !whole)) n_main = launch_wave(cl, dQ, layout, ld, chunk_begin, c_end, d_bits, u, s);
        else launch_tile<SrcConfigs, SinkBits>(im.f, im.wcfg, src, sink, u, (unsigned long long *)nullptr, grid, s);
        CK(cudaGetLastError());
        cl->stats.n_launches += n_main;
        if (c_end < n_cfg) return;
        src.n = n_cfg; src.s_begin = 0;
        launch_heavy<SrcConfigs, SinkBits>(im.f, src, sink, u, 148 * 16, s, (uint32_t *)cl->d_cnt.p + 7, cl);
        if (ev) CK(cudaEventRecord(ev[1], s));
        k_recheck_items<SrcConfigs, SinkBits><<<148 * 3, 128, 0, s>>>(im.d, src, sink, u);
        CK(cudaGetLastError());
        k_overflow_all<SrcConfigs, SinkBits><<<148 * 2, 128, 0, s>>>(im.d, src, sink, u);
        CK(cudaGetLastError());
        if (ev) CK(cudaEventRecord(ev[2], s));
        cl->stats.n_launches += 5;
        if (use_graph) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamEndCapture(s, &graph));
            aa_rx_cl::EdgeGraph g;
            memcpy(g.key, key, sizeof(key));
            CK(cudaGraphInstantiate(&g.exec, graph, 0));
            CK(cudaGraphDestroy(graph));
            g.used = ++cl->graph_clock;
            g.n_kernels = n_main + 5;
            if (cl->edge_graphs.size() >= 16) { /* replace the least recently used */
                size_t lru = 0;
                for (size_t k = 1; k < cl->edge_graphs.size(); k++)
                    if (cl->edge_graphs[k].used < cl->edge_graphs[lru].used) lru = k;
                cudaGraphExecDestroy(cl->edge_graphs[lru].exec);
                cl->edge_graphs[lru] = g;
            } else cl->edge_graphs.push_back(g);
            CK(cudaGraphLaunch(g.exec, s));
        }
        return;
    }
    if (chunk_begin != 0 || (chunk_end && chunk_end < n_cfg)) tmk_die("internal: chunked launch on the instrumented path");
    /* instrumented path (statistics / collision tracking): one thread per configuration */
    QSrc qs;
    qs.p = dQ; qs.p1 = nullptr; qs.layout = layout; qs.ld = ld;
    cl->d_unc.need(sizeof(uint32_t) * (n_cfg + 32));
    cl->d_cnt.need(64);
    CK(cudaMemsetAsync(cl->d_cnt.p, 0, 64, s));
    if (n_cfg == 0) return;
    cudaEvent_t *ev = timing_begin(cl, s);
    unsigned grid = (unsigned)((n_cfg + 127) / 128);
    k_check_cfg<float, false><<<grid, 128, 0, s>>>(im.f, qs, n_cfg, nullptr, nullptr, d_bits, (uint32_t *)cl->d_unc.p,
                                                   (uint32_t *)cl->d_cnt.p, pairhit, st);
    CK(cudaGetLastError());
    if (ev) CK(cudaEventRecord(ev[1], s));
    /* exact re-check of the uncertain configurations (device-side count, no host sync) */
    k_check_cfg<double, true><<<148 * 2, 128, 0, s>>>(im.d, qs, n_cfg, (const uint32_t *)cl->d_unc.p,
                                                      (const uint32_t *)cl->d_cnt.p, d_bits, nullptr, nullptr, pairhit, st);
    CK(cudaGetLastError());
    if (ev) CK(cudaEventRecord(ev[2], s));
    cl->stats.n_launches += 2;
}

static void launch_edges(struct aa_rx_cl *cl, size_t n_edge, const void *dQ0, const void *dQ1, int layout, size_t ld,
                         double seg_len, uint32_t *d_bits, int32_t *d_first, cudaStream_t s,
                         int n_levels = TMK_EDGE_LEVELS, size_t flat_states = 0) {
    ImageDev &im = cl->img;
    if (!(seg_len > 0)) tmk_die("edges: seg_len must be positive");
    if (n_edge > 0x7ffffff0u) tmk_die("edges: more than 2^31 edges in one call");
    DevStats *st = cl->stats_on ? (DevStats *)cl->d_stats.p : nullptr;
    if (flat_states && cl->use_tile && im.wcfg_flat.smem > 0 && !st && n_edge <= 4096 && !getenv("TMK_NO_FLAT")) {
        /* small batch (the caller bounded its states): every state of every edge in one tile launch, six launches in
         * all instead of the ~20 dependent ones of the level pipeline */
        UncList u = unc_list(cl, flat_states + 1024, s, true);
        const uint32_t ne = (uint32_t)n_edge;
        cl->d_nd.need(sizeof(int32_t) * (n_edge + 1));
        cl->d_efirst.need(sizeof(int32_t) * (n_edge + 1));
        cl->d_alive0.need(sizeof(uint32_t) * (n_edge + 2));
        int32_t *nd = (int32_t *)cl->d_nd.p, *first = (int32_t *)cl->d_efirst.p;
        uint32_t *off = (uint32_t *)cl->d_alive0.p, *cnt = (uint32_t *)cl->d_cnt.p;
        cudaEvent_t *ev = timing_begin(cl, s);
        k_edges_prep_flat<<<1, 1024, 0, s>>>(dQ0, dQ1, layout, ld, im.f.n_sub, ne, seg_len, nd, first, im.f.static_hit, off);
        SrcEdgeFlat src;
        src.lv.p0 = dQ0; src.lv.p1 = dQ1; src.lv.layout = layout; src.lv.ld = ld; src.lv.nd = nd; src.lv.n_edge = ne;
        src.lv.alive = nullptr; src.lv.n_alive = nullptr; src.lv.level = 0;
        src.off = off;
        SinkEdgeFirst sink;
        sink.first = first;
        const TileCfg &cfg = im.wcfg_flat;
        const unsigned grid = (unsigned)std::min<size_t>((flat_states + cfg.tile - 1) / cfg.tile + 1, (size_t)cfg.grid);
        launch_tile<SrcEdgeFlat, SinkEdgeFirst>(im.f, cfg, src, sink, u, (unsigned long long *)(cnt + 2), grid, s);
        k_heavy_mixed<SrcEdgeLevel, SinkEdgeFirst><<<148, 128, 0, s>>>(im.f, src.lv, sink, u);
        if (ev) CK(cudaEventRecord(ev[1], s));
        k_recheck_items<SrcEdgeLevel, SinkEdgeFirst><<<148 * 3, 128, 0, s>>>(im.d, src.lv, sink, u);
        k_edges_overflow_all<<<148 * 2, 128, 0, s>>>(im.d, src.lv, u, first);
        k_edges_finish<<<(ne + 255) / 256, 256, 0, s>>>(first, ne, d_bits, d_first);
        CK(cudaGetLastError());
        if (ev) CK(cudaEventRecord(ev[2], s));
        cl->stats.n_launches += 6;
        return;
    }
    const size_t small = getenv("TMK_EDGE_SMALL") ? (size_t)atol(getenv("TMK_EDGE_SMALL")) : 0; /* tuning aid */
    if (cl->use_tile && im.wcfg_e.smem > 0 && !st && n_edge >= small) {
        /* fast path: bisection levels through the tile pipeline (pipeline.cuh), replayed from a CUDA graph */
        UncList u = unc_list(cl, n_edge * 4, s, false);
        if (n_edge == 0) return;
        const uint32_t ne = (uint32_t)n_edge;
        cl->d_nd.need(sizeof(int32_t) * n_edge);
        cl->d_efirst.need(sizeof(int32_t) * n_edge);
        cl->d_alive0.need(sizeof(uint32_t) * n_edge);
        cl->d_alive1.need(sizeof(uint32_t) * n_edge);
        bool use_graph = !cl->timing && !getenv("TMK_NO_GRAPH") && !stream_is_capturing(s);
        uint64_t key[12] = {(uint64_t)n_edge | ((uint64_t)n_levels << 40), (uint64_t)(uintptr_t)dQ0, (uint64_t)(uintptr_t)dQ1, (uint64_t)layout * 1315423911u + ld,
                            0, (uint64_t)(uintptr_t)d_bits, (uint64_t)(uintptr_t)d_first, (uint64_t)(uintptr_t)s, im.serial,
                            (uint64_t)(uintptr_t)cl->d_unc.p ^ ((uint64_t)(uintptr_t)cl->d_mesh.p << 1) ^ ((uint64_t)(uintptr_t)cl->d_conv.p << 2),
                            (uint64_t)(uintptr_t)cl->d_nd.p ^ ((uint64_t)(uintptr_t)cl->d_efirst.p << 1) ^ ((uint64_t)(uintptr_t)cl->d_alive0.p << 2) ^
                                ((uint64_t)(uintptr_t)cl->d_alive1.p << 3),
                            (uint64_t)(uintptr_t)cl->d_spill.p ^ ((uint64_t)(uintptr_t)cl->d_tri.p << 1) ^ ((uint64_t)(uintptr_t)cl->d_cnt.p << 2)};
        memcpy(&key[4], &seg_len, sizeof(double));
        if (use_graph) {
            for (auto &g : cl->edge_graphs)
                if (0 == memcmp(g.key, key, sizeof(key))) {
                    g.used = ++cl->graph_clock;
                    CK(cudaGraphLaunch(g.exec, s));
                    cl->stats.n_launches += g.n_kernels; /* kernels run, not API calls */
                    return;
                }
            /* a graph is built for arguments seen before: a planner's batches differ call by call and
             * capture + instantiation costs more than enqueueing the launches once */
            bool seen = false;
            for (auto &k : cl->edge_seen) seen |= 0 == memcmp(k.data(), key, sizeof(key));
            if (!seen) {
                std::array<uint64_t, 12> k;
                memcpy(k.data(), key, sizeof(key));
                if (cl->edge_seen.size() >= 32) cl->edge_seen.erase(cl->edge_seen.begin());
                cl->edge_seen.push_back(k);
                use_graph = false;
            } else CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        }
        CK(cudaMemsetAsync(cl->d_cnt.p, 0, 512, s));
        /* d_cnt: [0] undecided count, [1] overflow, [2..3] evaluated states, [8 + L] alive count of level L */
        uint32_t *cnt = (uint32_t *)cl->d_cnt.p;
        int32_t *nd = (int32_t *)cl->d_nd.p, *first = (int32_t *)cl->d_efirst.p;
        cudaEvent_t *ev = timing_begin(cl, s);
        k_edges_prep<<<(ne + 255) / 256, 256, 0, s>>>(dQ0, dQ1, layout, ld, im.f.n_sub, ne, seg_len, nd, first,
                                                      im.f.static_hit, 1 << (n_levels - 1), cnt + 1);
        CK(cudaGetLastError());
        SrcEdgeLevel src;
        src.p0 = dQ0; src.p1 = dQ1; src.layout = layout; src.ld = ld; src.nd = nd; src.n_edge = ne;
        SinkEdgeFirst sink;
        sink.first = first;
        uint32_t *alive[2] = {(uint32_t *)cl->d_alive0.p, (uint32_t *)cl->d_alive1.p};
        const uint32_t *cur_alive = nullptr, *cur_n = nullptr;
        size_t launches = 1;
        /* heavy items (GJK, mesh traversals) of several levels are collected and run in one launch: a short
         * dependent kernel costs its latency, not its work, and a position decided late still merges
         * through atomicMin - the price is the few edges that bisect one level further than needed */
        const unsigned hmask = getenv("TMK_EDGE_HEAVY_MASK") ? (unsigned)strtoul(getenv("TMK_EDGE_HEAVY_MASK"), nullptr, 0) : 0x2222u;
        int group0 = 0;
        for (int L = 0; L < n_levels; L++) {
            src.alive = cur_alive; src.n_alive = cur_n; src.level = L;
            u.mesh_n = cnt + 32 + group0; u.conv_n = cnt + 48 + group0; u.tri_n = cnt + 64 + group0;
            launch_tile<SrcEdgeLevel, SinkEdgeFirst>(im.f, im.wcfg_e, src, sink, u, (unsigned long long *)(cnt + 2),
                                                     (unsigned)im.wcfg_e.grid, s);
            CK(cudaGetLastError());
            launches++;
            if (((hmask >> L) & 1) || L + 1 == n_levels) {
                k_heavy_mixed<SrcEdgeLevel, SinkEdgeFirst><<<148 * 4, 128, 0, s>>>(im.f, src, sink, u);
                CK(cudaGetLastError());
                launches++;
                group0 = L + 1;
            }
            if (L + 1 < n_levels) {
                uint32_t *out = alive[L & 1], *n_out = cnt + 8 + L;
                k_edges_compact<<<148 * 2, 256, 0, s>>>(cur_alive, cur_n, ne, nd, first, L, out, n_out);
                CK(cudaGetLastError());
                launches++;
                cur_alive = out; cur_n = n_out;
            }
        }
        if (ev) CK(cudaEventRecord(ev[1], s));
        /* undecided (state, pair) items of all levels, in double; positions merge through atomicMin */
        src.alive = nullptr; src.n_alive = nullptr; src.level = 0;
        k_recheck_items<SrcEdgeLevel, SinkEdgeFirst><<<148 * 3, 128, 0, s>>>(im.d, src, sink, u);
        CK(cudaGetLastError());
        k_edges_overflow_all<<<148 * 2, 128, 0, s>>>(im.d, src, u, first);
        CK(cudaGetLastError());
        k_edges_finish<<<(ne + 255) / 256, 256, 0, s>>>(first, ne, d_bits, d_first);
        CK(cudaGetLastError());
        if (ev) CK(cudaEventRecord(ev[2], s));
        cl->stats.n_launches += launches + 3;
        if (use_graph) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamEndCapture(s, &graph));
            aa_rx_cl::EdgeGraph g;
            memcpy(g.key, key, sizeof(key));
            CK(cudaGraphInstantiate(&g.exec, graph, 0));
            CK(cudaGraphDestroy(graph));
            g.used = ++cl->graph_clock;
            g.n_kernels = launches + 3;
            if (cl->edge_graphs.size() >= 16) { /* replace the least recently used */
                size_t lru = 0;
                for (size_t k = 1; k < cl->edge_graphs.size(); k++)
                    if (cl->edge_graphs[k].used < cl->edge_graphs[lru].used) lru = k;
                cudaGraphExecDestroy(cl->edge_graphs[lru].exec);
                cl->edge_graphs[lru] = g;
            } else cl->edge_graphs.push_back(g);
            CK(cudaGraphLaunch(g.exec, s));
        }
        return;
    }
    /* instrumented path: one warp per edge */
    QSrc qs;
    qs.p = dQ0; qs.p1 = dQ1; qs.layout = layout; qs.ld = ld;
    cl->d_unc.need(sizeof(uint32_t) * (n_edge + 32));
    cl->d_cnt.need(64);
    CK(cudaMemsetAsync(cl->d_cnt.p, 0, 64, s));
    if (n_edge == 0) return;
    cudaEvent_t *ev = timing_begin(cl, s);
    unsigned grid = (unsigned)((n_edge + 31) / 32);
    k_check_edges<float><<<grid, 256, 0, s>>>(im.f, qs, n_edge, seg_len, d_bits, d_first, (uint32_t *)cl->d_unc.p,
                                              (uint32_t *)cl->d_cnt.p, st);
    CK(cudaGetLastError());
    if (ev) CK(cudaEventRecord(ev[1], s));
    k_recheck_edges64<<<148 * 2, 128, 0, s>>>(im.d, qs, seg_len, (const uint32_t *)cl->d_unc.p,
                                              (const uint32_t *)cl->d_cnt.p, d_bits, d_first, st);
    CK(cudaGetLastError());
    if (ev) CK(cudaEventRecord(ev[2], s));
    cl->stats.n_launches += 2;
}

static void stats_end(struct aa_rx_cl *cl, size_t n_units, cudaStream_t s) {
    ImageDev &im = cl->img;
    CK(cudaMemcpyAsync(cl->h_cnt, cl->d_cnt.p, 16, cudaMemcpyDeviceToHost, s));
    DevStats hs;
    memset(&hs, 0, sizeof(hs));
    if (cl->stats_on) CK(cudaMemcpyAsync(&hs, cl->d_stats.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
    if (cl->track) {
        cl->last_pairhit.assign(im.h_pairs.size(), 0);
        if (!im.h_pairs.empty())
            CK(cudaMemcpyAsync(cl->last_pairhit.data(), cl->d_pairhit.p, im.h_pairs.size(), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    cl->stats.n_units = n_units;
    cl->stats.n_states = cl->stats_on ? hs.n_states : n_units;
    {
        unsigned long long evaluated = 0; /* the edge pipeline counts the states it evaluates */
        memcpy(&evaluated, cl->h_cnt + 2, sizeof(evaluated));
        if (!cl->stats_on && evaluated) cl->stats.n_states = evaluated;
    }
    cl->stats.n_joints = (uint64_t)im.f.n_joint;
    cl->stats.n_moving_geoms = (uint64_t)im.f.n_mg;
    cl->stats.n_pairs = (uint64_t)im.f.n_pair;
    for (int k = 0; k < 4; k++) cl->stats.n_narrow[k] = hs.n_narrow[k];
    cl->stats.n_uncertain = cl->h_cnt[0];
}

static size_t count_invalid(const uint32_t *bits, size_t n) {
    size_t valid = 0;
    const size_t full = n / 32; /* whole words */
    size_t w = 0;
    for (; w + 2 <= full; w += 2) { /* 64 bits at a time (the caller's buffer need not be 8-byte aligned) */
        uint64_t x;
        memcpy(&x, bits + w, sizeof(x));
        valid += (size_t)__builtin_popcountll(x);
    }
    for (; w < full; w++) valid += (size_t)__builtin_popcount(bits[w]);
    if (n & 31) valid += (size_t)__builtin_popcount(bits[full] & ((1u << (n & 31)) - 1));
    return n - valid;
}

/* The host-buffer call in two halves, so that a multi-GPU caller (multi.cu) can enqueue the slices of all devices
 * before it waits for any of them: begin() enqueues copies, kernels and the read-back of the validity words,
 * end() synchronises, collects the counters and counts the invalid units. */
static void batch_host_begin(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                             size_t n_q, const double *q_base, const void *Q, int layout, size_t ldQ,
                             uint32_t *valid_bits) {
    if (ldQ < n_sub) tmk_die("batch: ldQ < n_sub");
    HostSpan hs("batch_host_begin");
    ensure_image(cl, n_sub, sub_ids, n_q, q_base);
    stats_begin(cl);
    size_t esz = layout == TMK_Q_ROWS_F64 ? 8 : 4;
    size_t nw = (n_cfg + 31) / 32;
    cl->d_q.need(n_cfg * ldQ * esz + 16);
    cl->d_bits.need(sizeof(uint32_t) * (nw + 1));
    /* chunk = two full waves of the persistent grid (no tail effect inside a chunk) */
    size_t wave = (size_t)std::max(1, cl->img.wcfg.grid) * (size_t)std::max(32, cl->img.wcfg.tile);
    if (cl->img.f.grid.on) wave = std::max(wave, (size_t)262144); /* wavefront kernels: five launches per chunk */
    if (const char *e = getenv("TMK_HOST_WAVE")) wave = (size_t)std::max(4096L, atol(e)) & ~(size_t)1023; /* tuning aid */
    const size_t CH = 2 * wave;
    if (n_cfg <= CH + CH / 2 || cl->stats_on || cl->track || !cl->use_tile || cl->img.wcfg.smem <= 0) {
        if (n_cfg) CK(cudaMemcpyAsync(cl->d_q.p, Q, n_cfg * ldQ * esz, cudaMemcpyHostToDevice, cl->stream));
        launch_configs(cl, n_cfg, cl->d_q.p, layout, ldQ, (uint32_t *)cl->d_bits.p, cl->stream);
    } else {
        /* software pipeline: the copy of chunk k+1 (copy stream) overlaps the kernel of chunk k */
        size_t k = 0;
        for (size_t off = 0; off < n_cfg; k++) {
            /* two waves per chunk, one wave towards the end: what runs after the last copy is the last
             * chunk's kernel, so that chunk is kept short (the copy, not the kernels, sets the pace) */
            const size_t rem = n_cfg - off;
            size_t end = off + (rem > 3 * wave ? CH : wave);
            if (end + wave / 2 > n_cfg) end = n_cfg; /* fold a short tail into the last chunk */
            CK(cudaMemcpyAsync((char *)cl->d_q.p + off * ldQ * esz, (const char *)Q + off * ldQ * esz, (end - off) * ldQ * esz,
                               cudaMemcpyHostToDevice, cl->copy_stream));
            CK(cudaEventRecord(cl->ev_copy[k & 7], cl->copy_stream));
            CK(cudaStreamWaitEvent(cl->stream, cl->ev_copy[k & 7], 0));
            launch_configs(cl, n_cfg, cl->d_q.p, layout, ldQ, (uint32_t *)cl->d_bits.p, cl->stream, off, end);
            off = end;
        }
    }
    if (nw) CK(cudaMemcpyAsync(valid_bits, cl->d_bits.p, sizeof(uint32_t) * nw, cudaMemcpyDeviceToHost, cl->stream));
    hs.mark("enqueued");
}

static size_t batch_host_end(struct aa_rx_cl *cl, size_t n_units, const uint32_t *valid_bits) {
    HostSpan hs("batch_host_end");
    stats_end(cl, n_units, cl->stream);
    hs.mark("synchronised");
    size_t bad = count_invalid(valid_bits, n_units);
    cl->stats.n_invalid = bad;
    hs.mark("counted");
    return bad;
}

static size_t check_batch_host(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                               size_t n_q, const double *q_base, const void *Q, int layout, size_t ldQ,
                               uint32_t *valid_bits) {
    batch_host_begin(cl, n_cfg, n_sub, sub_ids, n_q, q_base, Q, layout, ldQ, valid_bits);
    return batch_host_end(cl, n_cfg, valid_bits);
}

extern "C" size_t aa_rx_cl_check_batch(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                       size_t n_q, const double *q_all_base, const double *Q_sub, size_t ldQ,
                                       uint32_t *valid_bits) {
    return check_batch_host(cl, n_cfg, n_sub, sub_ids, n_q, q_all_base, Q_sub, TMK_Q_ROWS_F64, ldQ, valid_bits);
}
extern "C" size_t aa_rx_cl_check_batch_f32(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub,
                                           const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                           const float *Q_sub, size_t ldQ, uint32_t *valid_bits) {
    return check_batch_host(cl, n_cfg, n_sub, sub_ids, n_q, q_all_base, Q_sub, TMK_Q_ROWS_F32, ldQ, valid_bits);
}

static void edges_host_begin(struct aa_rx_cl *cl, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                             size_t n_q, const double *q_all_base, const double *Q0, const double *Q1, size_t ldQ,
                             double seg_len, uint32_t *valid_bits, int32_t *first_invalid) {
    if (ldQ < n_sub) tmk_die("edges: ldQ < n_sub");
    ensure_image(cl, n_sub, sub_ids, n_q, q_all_base);
    stats_begin(cl);
    size_t nw = (n_edge + 31) / 32;
    cl->d_q.need(n_edge * ldQ * 8 + 16);
    cl->d_q1.need(n_edge * ldQ * 8 + 16);
    cl->d_bits.need(sizeof(uint32_t) * (nw + 1));
    if (first_invalid) cl->d_first.need(sizeof(int32_t) * (n_edge + 1));
    if (n_edge) {
        CK(cudaMemcpyAsync(cl->d_q.p, Q0, n_edge * ldQ * 8, cudaMemcpyHostToDevice, cl->stream));
        CK(cudaMemcpyAsync(cl->d_q1.p, Q1, n_edge * ldQ * 8, cudaMemcpyHostToDevice, cl->stream));
    }
    /* small batches (a planner's frontier): the longest edge bounds the bisection depth, and the dead levels
     * are launches the host can see and skip.  +2 segments of slack: the device count is authoritative (an edge
     * that needs more levels than were launched raises the overflow flag, see k_edges_prep). */
    int n_levels = TMK_EDGE_LEVELS;
    size_t flat_states = 0; /* > 0: an upper bound of the batch's states, small enough for the one-launch path */
    if (n_edge && n_edge <= 4096 && seg_len > 0) {
        double s2max = 0, states = 0;
        for (size_t e = 0; e < n_edge; e++) {
            double s2 = 0;
            for (size_t k = 0; k < n_sub; k++) {
                const double dq = Q1[e * ldQ + k] - Q0[e * ldQ + k];
                s2 += dq * dq;
            }
            if (s2 > s2max) s2max = s2;
            states += ceil(sqrt(s2) / seg_len) + 2; /* the device count is authoritative; this bounds the launch */
        }
        if (states <= 49152.0) flat_states = (size_t)states;
        const double nd_max = ceil(sqrt(s2max) / seg_len) + 2;
        if (nd_max < (double)(1 << (TMK_EDGE_LEVELS - 1))) {
            int L = 1; /* positions 2^(L-1) .. 2^L - 1 live at level L */
            while ((1 << (L - 1)) < (int)nd_max) L++;
            n_levels = std::min(TMK_EDGE_LEVELS, L);
        }
    }
    launch_edges(cl, n_edge, cl->d_q.p, cl->d_q1.p, TMK_Q_ROWS_F64, ldQ, seg_len, (uint32_t *)cl->d_bits.p,
                 first_invalid ? (int32_t *)cl->d_first.p : nullptr, cl->stream, n_levels, flat_states);
    if (nw) CK(cudaMemcpyAsync(valid_bits, cl->d_bits.p, sizeof(uint32_t) * nw, cudaMemcpyDeviceToHost, cl->stream));
    if (first_invalid && n_edge)
        CK(cudaMemcpyAsync(first_invalid, cl->d_first.p, sizeof(int32_t) * n_edge, cudaMemcpyDeviceToHost, cl->stream));
}

extern "C" size_t aa_rx_cl_check_edges(struct aa_rx_cl *cl, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                                       size_t n_q, const double *q_all_base, const double *Q0, const double *Q1,
                                       size_t ldQ, double seg_len, uint32_t *valid_bits, int32_t *first_invalid) {
    edges_host_begin(cl, n_edge, n_sub, sub_ids, n_q, q_all_base, Q0, Q1, ldQ, seg_len, valid_bits, first_invalid);
    return batch_host_end(cl, n_edge, valid_bits);
}

/* the two halves for multi.cu (declared in tmk_internal.h) */
extern "C" void tmk_cl_batch_host_begin(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                        size_t n_q, const double *q_base, const void *Q, int layout, size_t ldQ,
                                        uint32_t *valid_bits) {
    batch_host_begin(cl, n_cfg, n_sub, sub_ids, n_q, q_base, Q, layout, ldQ, valid_bits);
}
extern "C" void tmk_cl_edges_host_begin(struct aa_rx_cl *cl, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                                        size_t n_q, const double *q_base, const double *Q0, const double *Q1, size_t ldQ,
                                        double seg_len, uint32_t *valid_bits, int32_t *first_invalid) {
    edges_host_begin(cl, n_edge, n_sub, sub_ids, n_q, q_base, Q0, Q1, ldQ, seg_len, valid_bits, first_invalid);
}
extern "C" size_t tmk_cl_host_end(struct aa_rx_cl *cl, size_t n_units, const uint32_t *valid_bits) {
    return batch_host_end(cl, n_units, valid_bits);
}
extern "C" int tmk_cl_device(const struct aa_rx_cl *cl) { return cl->dev; }
extern "C" void *tmk_cl_stream(const struct aa_rx_cl *cl) { return (void *)cl->stream; }
extern "C" const struct aa_rx_sg *tmk_cl_sg(const struct aa_rx_cl *cl) { return cl->sg; }

extern "C" void aa_rx_cl_check_batch_dev(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                         size_t n_q, const double *q_all_base, const void *dQ, int layout, size_t ld,
                                         uint32_t *d_valid_bits, void *stream) {
    ensure_image(cl, n_sub, sub_ids, n_q, q_all_base);
    cudaStream_t s = stream ? (cudaStream_t)stream : cl->stream;
    if (cl->stats_on) { cl->d_stats.need(sizeof(DevStats)); CK(cudaMemsetAsync(cl->d_stats.p, 0, sizeof(DevStats), s)); }
    cl->stats.n_launches = 0;
    launch_configs(cl, n_cfg, dQ, layout, ld, d_valid_bits, s);
}

extern "C" void aa_rx_cl_check_edges_dev(struct aa_rx_cl *cl, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                                         size_t n_q, const double *q_all_base, const void *dQ0, const void *dQ1,
                                         int layout, size_t ld, double seg_len, uint32_t *d_valid_bits,
                                         int32_t *d_first_invalid, void *stream) {
    ensure_image(cl, n_sub, sub_ids, n_q, q_all_base);
    cudaStream_t s = stream ? (cudaStream_t)stream : cl->stream;
    if (cl->stats_on) { cl->d_stats.need(sizeof(DevStats)); CK(cudaMemsetAsync(cl->d_stats.p, 0, sizeof(DevStats), s)); }
    cl->stats.n_launches = 0;
    launch_edges(cl, n_edge, dQ0, dQ1, layout, ld, seg_len, d_valid_bits, d_first_invalid, s);
}

extern "C" void aa_rx_cl_sync(struct aa_rx_cl *cl, void *stream) {
    CK(cudaStreamSynchronize(stream ? (cudaStream_t)stream : cl->stream));
}

extern "C" void aa_rx_cl_track_collisions(struct aa_rx_cl *cl, int enable) { cl->track = enable ? 1 : 0; }

extern "C" void aa_rx_cl_batch_collisions(struct aa_rx_cl *cl, struct aa_rx_cl_set *set) {
    if (set->n_f != cl->n_f) tmk_die("aa_rx_cl_batch_collisions: size mismatch");
    ImageDev &im = cl->img;
    for (size_t p = 0; p < cl->last_pairhit.size() && p < im.pair_fa.size(); p++)
        if (cl->last_pairhit[p]) aa_rx_cl_set_set(set, im.pair_fa[p], im.pair_fb[p], 1);
}

/* ------------------------------------------------------------------ FP32 roofline probe
 * The denominator of the FP32 roofline: 8 independent FFMA chains per thread, 8 CTAs of 256 threads
 * per SM, long enough to reach steady clocks.  Measured, not derived. */
__global__ void __launch_bounds__(256) k_fma_probe(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[0] = s; /* never true: keeps the chain alive */
}

extern "C" double tmk_fp32_peak_tflops(void) {
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *d = nullptr;
    CK(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 4096, grid = sms * 8;
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0, 0));
        k_fma_probe<<<grid, 256>>>(d, iters, 0.999f, 0.001f);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 64.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

#ifdef TMK_PHASE_CLOCKS
extern "C" void tmk_phase_cycles(unsigned long long *out, int reset) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, tmk::g_phase_cycles, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0};
        CK(cudaMemcpyToSymbol(tmk::g_phase_cycles, z, sizeof(z)));
    }
}
#endif

extern "C" void aa_rx_cl_kernel_timing(struct aa_rx_cl *cl, int enable) {
    cl->timing = enable ? 1 : 0;
    cl->ev_used = 0;
}

extern "C" size_t aa_rx_cl_kernel_time(struct aa_rx_cl *cl, double *ms_main, double *ms_recheck) {
    size_t n = cl->ev_used / 3;
    double a = 0, b = 0;
    for (size_t k = 0; k < n; k++) {
        cudaEvent_t *e = &cl->ev_pool[3 * k];
        CK(cudaEventSynchronize(e[2]));
        float t0 = 0, t1 = 0;
        CK(cudaEventElapsedTime(&t0, e[0], e[1]));
        CK(cudaEventElapsedTime(&t1, e[1], e[2]));
        a += t0; b += t1;
    }
    if (ms_main) *ms_main = a;
    if (ms_recheck) *ms_recheck = b;
    cl->ev_used = 0;
    return n;
}

extern "C" void aa_rx_cl_stats_enable(struct aa_rx_cl *cl, int enable) { cl->stats_on = enable ? 1 : 0; }
extern "C" void aa_rx_cl_batch_stats(struct aa_rx_cl *cl, struct tmk_batch_stats *out) { *out = cl->stats; }

/* ------------------------------------------------------------------ K1 standalone: batched FK */

static FkImage *fk_image(const struct aa_rx_sg *sg, size_t n_sub, const aa_rx_config_id *sub_ids, size_t n_q,
                         const double *q_base, size_t n_frames, const aa_rx_frame_id *frames) {
    if (!sg || !sg->inited) tmk_die("aa_rx_sg_tf_batch: scene graph not initialised");
    if (n_q != sg->n_q) tmk_die("aa_rx_sg_tf_batch: n_q mismatch");
    if (n_sub > TMK_MAXSUB) tmk_die("aa_rx_sg_tf_batch: n_sub = %zu > %d", n_sub, TMK_MAXSUB);
    SgDev *sd = sg_dev(sg);
    if (!sd->fk) sd->fk = new FkImage();
    FkImage *im = sd->fk;
    const size_t n_f = sg->n_f;
    const bool all = frames == nullptr;
    bool same = im->valid && im->sg_serial == sg->serial && im->all_frames == all && im->sub.size() == n_sub &&
                0 == memcmp(im->sub.data(), sub_ids, sizeof(int) * n_sub) && 0 == memcmp(im->q_base.data(), q_base, sizeof(double) * n_q) &&
                (all || (im->frames.size() == n_frames && 0 == memcmp(im->frames.data(), frames, sizeof(int) * n_frames)));
    if (same) return im;
    for (size_t k = 0; k < n_sub; k++)
        if (sub_ids[k] < 0 || (size_t)sub_ids[k] >= n_q) tmk_die("aa_rx_sg_tf_batch: bad configuration id %d", sub_ids[k]);
    std::vector<double> rel(7 * n_f), ab(7 * n_f);
    tmk_dev_fk64(sg, q_base, rel.data(), ab.data());
    /* fold the fixed (and the not-varied) frames into the varied joint they hang on: the collision image's rule */
    std::vector<int> slot(n_f, -1);
    std::vector<double> fold(7 * n_f, 0.0);
    std::vector<DJoint<float>> joints;
    for (size_t i = 0; i < n_f; i++) {
        const int p = sg->parent[i];
        int qs = -1;
        if (sg->type[i] != TMK_FIXED)
            for (size_t k = 0; k < n_sub; k++)
                if (sub_ids[k] == sg->qidx[i]) qs = (int)k;
        if (qs >= 0) {
            DJoint<double> J;
            memset(&J, 0, sizeof(J));
            const double *E = sg->E + 7 * i;
            if (p < 0) { memcpy(J.E, E, sizeof(double) * 7); J.parent = -1; }
            else if (slot[p] < 0) { tmk_tfmul(&ab[7 * p], E, J.E); J.parent = -1; }
            else { tmk_tfmul(&fold[7 * p], E, J.E); J.parent = slot[p]; }
            memcpy(J.ax, sg->axis + 3 * i, sizeof(double) * 3);
            J.off = sg->offset[i]; J.type = sg->type[i]; J.qslot = qs;
            slot[i] = (int)joints.size();
            DJoint<float> Jf;
            cvt_joint(J, Jf);
            joints.push_back(Jf);
            const double id7[7] = {0, 0, 0, 1, 0, 0, 0};
            memcpy(&fold[7 * i], id7, sizeof(id7));
        } else if (p >= 0 && slot[p] >= 0) {
            slot[i] = slot[p];
            tmk_tfmul(&fold[7 * p], &rel[7 * i], &fold[7 * i]);
        }
    }
    if (joints.size() > TMK_MAXJ) tmk_die("aa_rx_sg_tf_batch: %zu varied joints > %d", joints.size(), TMK_MAXJ);
    const size_t n_out = all ? n_f : n_frames;
    std::vector<FkOut> outs(n_out);
    for (size_t k = 0; k < n_out; k++) {
        const int f = all ? (int)k : frames[k];
        if (f < 0 || (size_t)f >= n_f) tmk_die("aa_rx_sg_tf_batch: bad frame id %d", f);
        const double *src = slot[f] >= 0 ? &fold[7 * (size_t)f] : &ab[7 * (size_t)f];
        for (int i = 0; i < 7; i++) outs[k].tf[i] = (float)src[i];
        outs[k].slot = slot[f];
        outs[k].out = (int)k;
    }
    /* the kernel writes a joint's frames while its pose is in registers: table sorted by slot (constants first);
     * joints whose children do not follow them directly are the only ones it keeps in its per-thread array */
    std::stable_sort(outs.begin(), outs.end(), [](const FkOut &x, const FkOut &y) { return x.slot < y.slot; });
    for (size_t j = 0; j < joints.size(); j++) joints[j].pad = 0;
    for (size_t j = 0; j < joints.size(); j++)
        if (joints[j].parent >= 0 && joints[j].parent != (int)j - 1) joints[joints[j].parent].pad = 1;
    CK(cudaDeviceSynchronize()); /* launches in flight may still read the old tables */
    upload(im->joints, joints, 0);
    upload(im->outs, outs, 0);
    CK(cudaDeviceSynchronize());
    im->n_joint = (int)joints.size(); im->n_out = (int)n_out;
    im->sub.assign(sub_ids, sub_ids + n_sub);
    im->q_base.assign(q_base, q_base + n_q);
    im->frames.clear();
    if (!all) im->frames.assign(frames, frames + n_frames);
    im->all_frames = all; im->sg_serial = sg->serial; im->valid = true;
    return im;
}

extern "C" void aa_rx_sg_tf_batch_dev(const struct aa_rx_sg *sg, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                      size_t n_q, const double *q_all_base, const void *dQ, int layout, size_t ld,
                                      size_t n_frames, const aa_rx_frame_id *frames, void *dTF, int tf_layout, size_t ld_tf,
                                      void *stream) {
    if (layout < 0 || layout > 2) tmk_die("aa_rx_sg_tf_batch_dev: unknown layout %d", layout);
    if (tf_layout == TMK_TF_SOA_F32 && ld_tf < n_cfg) tmk_die("aa_rx_sg_tf_batch_dev: ld_tf < n_cfg");
    FkImage *im = fk_image(sg, n_sub, sub_ids, n_q, q_all_base, n_frames, frames);
    if (n_cfg == 0 || im->n_out == 0) return;
    const size_t smem = (((size_t)im->n_joint * sizeof(DJoint<float>) + 15) & ~(size_t)15) + (size_t)im->n_out * sizeof(FkOut);
    if (smem > 200 * 1024) tmk_die("aa_rx_sg_tf_batch: %d frames exceed the kernel's frame table", im->n_out);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n_cfg + 255) / 256);
    if (tf_layout == TMK_TF_SOA_F32) {
        static bool attr = false;
        if (!attr) { CK(cudaFuncSetAttribute(k_fk_frames<TMK_TF_SOA_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
        k_fk_frames<TMK_TF_SOA_F32><<<grid, 256, smem, s>>>((const DJoint<float> *)im->joints.p, im->n_joint, (const FkOut *)im->outs.p, im->n_out,
                                                         dQ, layout, ld, n_cfg, dTF, ld_tf);
    } else if (tf_layout == TMK_TF_ROWS_F64) {
        static bool attr = false;
        if (!attr) { CK(cudaFuncSetAttribute(k_fk_frames<TMK_TF_ROWS_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
        k_fk_frames<TMK_TF_ROWS_F64><<<grid, 256, smem, s>>>((const DJoint<float> *)im->joints.p, im->n_joint, (const FkOut *)im->outs.p, im->n_out,
                                                          dQ, layout, ld, n_cfg, dTF, ld_tf);
    } else tmk_die("aa_rx_sg_tf_batch_dev: unknown output layout %d", tf_layout);
    CK(cudaGetLastError());
}

extern "C" void aa_rx_sg_tf_batch(const struct aa_rx_sg *sg, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                  size_t n_q, const double *q_all_base, const double *Q_sub, size_t ldQ,
                                  size_t n_frames, const aa_rx_frame_id *frames, double *TF_abs_out) {
    if (ldQ < n_sub) tmk_die("aa_rx_sg_tf_batch: ldQ < n_sub");
    FkImage *im = fk_image(sg, n_sub, sub_ids, n_q, q_all_base, n_frames, frames);
    if (n_cfg == 0 || im->n_out == 0) return;
    SgDev *sd = sg_dev(sg);
    DBuf &dq = sd->b_q, &dout = sd->b_out;
    dq.need(n_cfg * ldQ * 8);
    dout.need(n_cfg * (size_t)im->n_out * 7 * 8);
    CK(cudaMemcpyAsync(dq.p, Q_sub, n_cfg * ldQ * 8, cudaMemcpyHostToDevice, 0));
    aa_rx_sg_tf_batch_dev(sg, n_cfg, n_sub, sub_ids, n_q, q_all_base, dq.p, TMK_Q_ROWS_F64, ldQ, n_frames, frames, dout.p,
                          TMK_TF_ROWS_F64, 0, nullptr);
    CK(cudaMemcpy(TF_abs_out, dout.p, n_cfg * (size_t)im->n_out * 7 * 8, cudaMemcpyDeviceToHost));
}
