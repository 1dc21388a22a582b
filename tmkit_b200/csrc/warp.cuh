/* warp.cuh - the batched validity kernel, second generation: warp-autonomous pipelines.
 *
 * One CTA = 8 warps; every warp owns groups of 32 states (one state per lane) and runs the whole
 * pipeline for them with no CTA-wide barrier after the scene image has been staged:
 *
 *   stage   scene tables (joints, geometry, grouped pair list, quick tables) copied once per CTA into
 *           shared memory with TMA 1-D bulk copies completing on an mbarrier (images too large
 *           for the staging budget are read through L1 instead)
 *   P1+P2   one lane = one state.  The joint chain runs in registers; as soon as the pose of a
 *           moving geometry exists it is written to the warp's pose slab ([geom*7+k][lane], bank
 *           conflict free) and its partner list is walked.  The list is warp-uniform, so the
 *           bounding-sphere test and the QUICK CERTIFICATES (closed forms for spheres, face /
 *           axis separating-axis tests and centre-containment tests for boxes and cylinders) run
 *           without divergence.  Pairs they cannot decide are compacted with ballot/popc into the
 *           warp's queue.
 *   P3      one lane = one queued (state, pair): SAT / GJK / BVH narrowphase from dev_math.cuh on
 *           the poses in the slab.
 *   out     a state is valid unless a HIT was certified.  Pairs the FP32 certificates cannot decide
 *           (UNC) are appended to a global (state, pair) list; k_recheck_items re-evaluates exactly
 *           those pairs in double precision and clears the state's result on a HIT.
 *
 * Src supplies the states (configurations, or interior states of swept edges); Sink consumes the
 * verdicts (validity bitmap, or per-edge first-invalid positions).
 */
#pragma once
#include "kernels.cuh"

namespace tmk {

#define TMK_WARPS 8
#define TMK_WQ 384 /* queue entries per warp */

/* world-frame quick geometry: centre, axes (columns of the rotation), dims, bounding radius */
struct SQuick {
    float cx, cy, cz, brad;
    float a0[3], a1[3], a2[3];
    float h0, h1, h2;
};

struct WarpCfg {
    int stage; /* 1: image staged in shared memory */
    int off_joint, off_mg, off_sg, off_pairs, off_sq, off_jg, off_pa, off_warp;
    int bytes_joint, bytes_mg, bytes_sg, bytes_pairs, bytes_sq, bytes_jg, bytes_pa;
    int warp_bytes, woff_q, woff_queue, woff_flags, woff_unc, woff_jslab;
    int smem, grid;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ QG qg_from_pose(const DGeom<float> &G, const Xf<float> &X) {
    return qg_make(G.shape, G.dim[0], G.dim[1], G.dim[2], G.brad, X, mk<float>(G.dim[0], G.dim[1], G.dim[2]));
}
__device__ __forceinline__ QG qg_from_static(const SQuick &s) {
    QG g;
    g.c = mk<float>(s.cx, s.cy, s.cz);
    g.x = mk<float>(s.a0[0], s.a0[1], s.a0[2]);
    g.y = mk<float>(s.a1[0], s.a1[1], s.a1[2]);
    g.z = mk<float>(s.a2[0], s.a2[1], s.a2[2]);
    g.h0 = s.h0; g.h1 = s.h1; g.h2 = s.h2;
    g.brad = s.brad;
    return g;
}

__device__ __forceinline__ void slab_store(float *slab, int g, int lane, const Xf<float> &X) {
    float *o = slab + (size_t)g * 7 * 32 + lane;
    o[0] = X.r.x; o[32] = X.r.y; o[64] = X.r.z; o[96] = X.r.w;
    o[128] = X.v.x; o[160] = X.v.y; o[192] = X.v.z;
}
__device__ __forceinline__ Xf<float> slab_load(const float *slab, int g, int lane) {
    const float *o = slab + (size_t)g * 7 * 32 + lane;
    Xf<float> X;
    X.r.x = o[0]; X.r.y = o[32]; X.r.z = o[64]; X.r.w = o[96];
    X.v.x = o[128]; X.v.y = o[160]; X.v.z = o[192];
    return X;
}

struct UncList {
    uint2 *items;       /* (state, pair) */
    uint32_t *n;        /* appended count (may exceed cap) */
    uint32_t cap;
    uint32_t *overflow; /* set when an item was dropped: the brute-force double kernel then runs */
};
__device__ __forceinline__ void unc_emit(const UncList &u, uint2 item) {
    uint32_t i = atomicAdd(u.n, 1u);
    if (i < u.cap) u.items[i] = item;
    else *u.overflow = 1u;
}

/* ------------------------------------------------------------------ sources of states, sinks of verdicts
 * Src:  count(), prepare(s) -> State (with .live), load1<T>(State, k) = k-th joint value,
 *       unc_item(State, pair) / from_unc(item, &pair) to carry a state through the undecided list.
 * Sink: emit(State, g, hit, lane) called convergently by the 32 lanes of group g; mark_hit(State). */
struct SrcConfigs {
    const void *p;
    int layout;
    size_t ld, n;
    struct State { size_t s; bool live; };
    __device__ __forceinline__ size_t count() const { return n; }
    __device__ __forceinline__ State prepare(size_t s) const { State st; st.s = s; st.live = s < n; return st; }
    template <typename T> __device__ __forceinline__ T load1(const State &st, int k) const {
        return load_q<T>(p, layout, ld, st.s, k);
    }
    __device__ __forceinline__ uint2 unc_item(const State &st, uint32_t pair) const { return make_uint2((uint32_t)st.s, pair); }
    __device__ __forceinline__ State from_unc(uint2 it, uint32_t *pair) const { *pair = it.y; return prepare(it.x); }
};

struct SinkBits {
    uint32_t *bits;
    __device__ __forceinline__ void emit(const SrcConfigs::State &st, size_t g, bool hit, int lane) const {
        (void)st;
        const unsigned valid = __ballot_sync(0xffffffffu, !hit);
        if (lane == 0) bits[g] = valid;
    }
    __device__ __forceinline__ void mark_hit(const SrcConfigs::State &st) const {
        atomicAnd(&bits[st.s >> 5], ~(1u << (st.s & 31)));
    }
};

/* Swept edges, one bisection level at a time (OMPL DiscreteMotionValidator order).
 * Level 0 is the end state (position 0 in test order); level L >= 1 holds the interior positions
 * 2^(L-1) .. 2^L - 1 (breadth-first order of the bisection tree).  An edge is alive at level L
 * if no state of an earlier level was certified invalid.  first[e] = min invalid position. */
#define TMK_EDGE_LEVELS 14
#define TMK_FIRST_NONE 0x7fffffff
struct SrcEdgeLevel {
    const void *p0, *p1;
    int layout;
    size_t ld;
    const int32_t *nd;       /* segments per edge */
    const uint32_t *alive;   /* edges alive at this level (NULL: all n_edge edges) */
    const uint32_t *n_alive; /* device count of `alive` (NULL with alive == NULL) */
    uint32_t n_edge;
    int level;
    struct State { uint32_t e; int pos, nd; double t; bool live; };
    __device__ __forceinline__ size_t count() const {
        size_t na = alive ? (size_t)*n_alive : (size_t)n_edge;
        return level <= 1 ? na : na << (level - 1);
    }
    __device__ __forceinline__ State at(uint32_t e, int pos) const {
        State st;
        st.e = e; st.pos = pos; st.nd = nd[e];
        st.live = pos < st.nd; /* positions 0 .. nd-1 exist */
        st.t = 1.0;
        if (st.live && pos > 0) st.t = __ddiv_rn((double)ompl_mid(st.nd, pos - 1), (double)st.nd);
        return st;
    }
    __device__ __forceinline__ State prepare(size_t s) const {
        const int sh = level <= 1 ? 0 : level - 1;
        const size_t ai = s >> sh;
        const size_t na = alive ? (size_t)*n_alive : (size_t)n_edge;
        if (ai >= na) { State st; st.e = 0; st.pos = 0; st.nd = 0; st.t = 1.0; st.live = false; return st; }
        const uint32_t e = alive ? alive[ai] : (uint32_t)ai;
        const int pos = level == 0 ? 0 : (1 << sh) + (int)(s & ((size_t(1) << sh) - 1));
        return at(e, pos);
    }
    template <typename T> __device__ __forceinline__ T load1(const State &st, int k) const {
        const double b = load_q<double>(p1, layout, ld, st.e, k);
        if (st.pos == 0) return (T)b;
        const double a = load_q<double>(p0, layout, ld, st.e, k);
        return (T)__dadd_rn(a, __dmul_rn(__dsub_rn(b, a), st.t));
    }
    /* undecided item: x = edge, y = position << 19 | pair */
    __device__ __forceinline__ uint2 unc_item(const State &st, uint32_t pair) const {
        return make_uint2(st.e, ((uint32_t)st.pos << 19) | pair);
    }
    __device__ __forceinline__ State from_unc(uint2 it, uint32_t *pair) const {
        *pair = it.y & 0x7ffffu;
        return at(it.x, (int)(it.y >> 19));
    }
};

struct SinkEdgeFirst {
    int32_t *first; /* min invalid position per edge, TMK_FIRST_NONE = none so far */
    __device__ __forceinline__ void emit(const SrcEdgeLevel::State &st, size_t g, bool hit, int lane) const {
        (void)g; (void)lane;
        if (st.live && hit) atomicMin(&first[st.e], st.pos);
    }
    __device__ __forceinline__ void mark_hit(const SrcEdgeLevel::State &st) const {
        if (st.live) atomicMin(&first[st.e], st.pos);
    }
};

/* ------------------------------------------------------------------ the kernel */
template <class Src, class Sink>
__global__ void __launch_bounds__(TMK_WARPS * 32, 2)
k_states_warp(Image<float> im, WarpCfg cfg, Src src, Sink sink, UncList unc, unsigned long long *n_states) {
    const size_t n = src.count();
    if (n == 0) return; /* dead bisection levels cost a launch, not a staging */
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DJoint<float> *J = im.joints;
    const DGeom<float> *MG = im.mg;
    const DGeom<float> *SG = im.sg;
    const uint32_t *PAIRS = im.pairs;
    const SQuick *SQ = (const SQuick *)im.sq;
    const int *JG = im.jg_begin, *PA = im.pa_begin;

    if (cfg.stage) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t total = (uint32_t)(cfg.bytes_joint + cfg.bytes_mg + cfg.bytes_sg + cfg.bytes_pairs + cfg.bytes_sq +
                                        cfg.bytes_jg + cfg.bytes_pa);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(total)
                         : "memory");
            bulk_g2s(smem + cfg.off_joint, im.joints, (uint32_t)cfg.bytes_joint, &bar);
            bulk_g2s(smem + cfg.off_mg, im.mg, (uint32_t)cfg.bytes_mg, &bar);
            if (cfg.bytes_sg) bulk_g2s(smem + cfg.off_sg, im.sg, (uint32_t)cfg.bytes_sg, &bar);
            if (cfg.bytes_pairs) bulk_g2s(smem + cfg.off_pairs, im.pairs, (uint32_t)cfg.bytes_pairs, &bar);
            if (cfg.bytes_sq) bulk_g2s(smem + cfg.off_sq, im.sq, (uint32_t)cfg.bytes_sq, &bar);
            bulk_g2s(smem + cfg.off_jg, im.jg_begin, (uint32_t)cfg.bytes_jg, &bar);
            bulk_g2s(smem + cfg.off_pa, im.pa_begin, (uint32_t)cfg.bytes_pa, &bar);
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0u)
                         : "memory");
        }
        J = (const DJoint<float> *)(smem + cfg.off_joint);
        MG = (const DGeom<float> *)(smem + cfg.off_mg);
        SG = (const DGeom<float> *)(smem + cfg.off_sg);
        PAIRS = (const uint32_t *)(smem + cfg.off_pairs);
        SQ = (const SQuick *)(smem + cfg.off_sq);
        JG = (const int *)(smem + cfg.off_jg);
        PA = (const int *)(smem + cfg.off_pa);
    }

    unsigned char *wbase = smem + cfg.off_warp + (size_t)warp * cfg.warp_bytes;
    float *slab = (float *)wbase;                         /* [n_mg*7][32] geometry poses */
    float *qs = (float *)(wbase + cfg.woff_q);            /* [n_sub][32] joint values */
    uint32_t *queue = (uint32_t *)(wbase + cfg.woff_queue);
    volatile uint8_t *flags = wbase + cfg.woff_flags;     /* [32] HIT flags, one per lane's state */
    float *jslab = (float *)(wbase + cfg.woff_jslab);     /* poses of joints with non-consecutive children */
    uint32_t *unc_x = (uint32_t *)(wbase + cfg.woff_unc);  /* [32] + [32]: undecided-item words of each lane's state */
    uint32_t *unc_y = unc_x + 32;

    const size_t n_groups = (n + 31) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_joint = im.n_joint, n_sub = im.n_sub;

    auto flush = [&](uint32_t qn, const typename Src::State &st) {
        __syncwarp();
        for (uint32_t i = lane; i < qn; i += 32) {
            const uint32_t e = queue[i];
            const int t = e >> 27, p = e & 0x7ffffff;
            if (flags[t]) continue;
            const uint32_t w = PAIRS[p];
            const int a = w & 0xff, b = (w >> 8) & 0xffff;
            const Xf<float> XA = slab_load(slab, a, t);
            const DGeom<float> *B;
            Xf<float> XB;
            if ((w >> 24) & 1) {
                B = &SG[b];
                XB = xload(B->tf);
            } else {
                B = &MG[b];
                XB = slab_load(slab, b, t);
            }
            const int r = test_geoms(MG[a], XA, *B, XB, im.meshes);
            if (r == HIT) flags[t] = 1;
            else if (r == UNC) {
                /* the item belongs to lane t's state: fetch its identity from that lane's registers is
                 * not possible here, so the owner lane's unc item words were parked in shared memory */
                unc_emit(unc, make_uint2(unc_x[t], unc_y[t] | (uint32_t)p));
            }
        }
        __syncwarp();
        (void)st;
    };

    for (size_t g = (size_t)blockIdx.x * TMK_WARPS + warp; g < n_groups; g += (size_t)gridDim.x * TMK_WARPS) {
        const size_t s = g * 32 + lane;
        const typename Src::State st = src.prepare(s);
        const bool live = st.live && !im.static_hit;
        bool hit = st.live && im.static_hit;
        flags[lane] = live ? 0 : 1; /* dead lanes: nothing to evaluate, nothing queued */
        {
            const uint2 it = src.unc_item(st, 0u);
            unc_x[lane] = it.x; unc_y[lane] = it.y;
        }
        if (st.live)
            for (int k = 0; k < n_sub; k++) qs[k * 32 + lane] = src.template load1<float>(st, k);
        if (n_states) {
            const unsigned lm = __ballot_sync(0xffffffffu, live);
            if (lane == 0 && lm) atomicAdd(n_states, (unsigned long long)__popc(lm));
        }
        __syncwarp();
        uint32_t qn = 0;
        Xf<float> cur;
        cur.r.x = cur.r.y = cur.r.z = 0.f; cur.r.w = 1.f; cur.v = mk<float>(0.f, 0.f, 0.f);
        for (int j = 0; j < n_joint; j++) {
            const DJoint<float> &Jt = J[j];
            const int par_id = Jt.parent;
            Xf<float> par = cur;
            if (par_id >= 0 && par_id != j - 1) par = slab_load(jslab, J[par_id].pad - 1, lane);
            cur = fk_joint(xload(Jt.E), mk<float>(Jt.ax[0], Jt.ax[1], Jt.ax[2]), Jt.off, Jt.type, qs[Jt.qslot * 32 + lane],
                           par_id >= 0 ? &par : (const Xf<float> *)nullptr);
            if (Jt.pad > 0) slab_store(jslab, Jt.pad - 1, lane, cur);
            for (int a = JG[j]; a < JG[j + 1]; a++) {
                const DGeom<float> &A = MG[a];
                const Xf<float> XA = xmul(cur, xload(A.tf));
                slab_store(slab, a, lane, XA);
                const QG qa = qg_from_pose(A, XA);
                const int ka = A.shape;
                const int p_end = PA[a + 1];
                for (int p = PA[a]; p < p_end; p++) {
                    const uint32_t w = PAIRS[p];
                    const int b = (w >> 8) & 0xffff;
                    bool need = false;
                    if (live && !hit) {
                        int r;
                        if ((w >> 24) & 1) {
                            r = quick_pair(ka, qa, SG[b].shape, qg_from_static(SQ[b]));
                        } else {
                            const DGeom<float> &B = MG[b];
                            r = quick_pair(ka, qa, B.shape, qg_from_pose(B, slab_load(slab, b, lane)));
                        }
                        if (r == HIT) { hit = true; flags[lane] = 1; }
                        else if (r == UNC) unc_emit(unc, src.unc_item(st, (uint32_t)p));
                        else if (r == NEED) need = true;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, need);
                    if (m) {
                        if (qn + 32 > TMK_WQ) { flush(qn, st); qn = 0; }
                        if (need) queue[qn + __popc(m & lt_mask)] = ((uint32_t)lane << 27) | (uint32_t)p;
                        qn += __popc(m);
                    }
                }
            }
        }
        flush(qn, st);
        hit = hit || (live && flags[lane] != 0);
        sink.emit(st, g, hit || !st.live, lane);
        __syncwarp();
    }
}

/* exact re-check of the (state, pair) items the FP32 certificates could not decide */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_recheck_items(Image<double> im, Src src, Sink sink, UncList unc) {
    uint32_t n = *unc.n;
    if (n > unc.cap) n = unc.cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t pair;
        const typename Src::State st = src.from_unc(unc.items[i], &pair);
        if (!st.live) continue;
        double q[TMK_MAXSUB];
        for (int k = 0; k < im.n_sub; k++) q[k] = src.template load1<double>(st, k);
        Xf<double> gp[TMK_MAXMG];
        fk_config(im, q, gp);
        const uint32_t w = im.pairs[pair];
        const int a = w & 0xff, b = (w >> 8) & 0xffff;
        const DGeom<double> &A = im.mg[a];
        const DGeom<double> &B = (w >> 24) & 1 ? im.sg[b] : im.mg[b];
        const Xf<double> XB = (w >> 24) & 1 ? xload(B.tf) : gp[b];
        if (test_geoms(A, gp[a], B, XB, im.meshes) == HIT) sink.mark_hit(st);
    }
}

/* safety net: only does work when the undecided list overflowed - every state in double */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_overflow_all(Image<double> im, Src src, Sink sink, UncList unc) {
    if (*unc.overflow == 0) return;
    const size_t n = src.count();
    for (size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x) {
        const typename Src::State st = src.prepare(s);
        if (!st.live) continue;
        double q[TMK_MAXSUB];
        for (int k = 0; k < im.n_sub; k++) q[k] = src.template load1<double>(st, k);
        if (eval_config(im, q, (uint8_t *)nullptr, (DevStats *)nullptr) == HIT) sink.mark_hit(st);
    }
}

/* ------------------------------------------------------------------ swept-edge bookkeeping kernels */
/* segments per edge (double, bit-identical to the host restatement) and first[] reset */
__global__ void k_edges_prep(const void *p0, const void *p1, int layout, size_t ld, int n_sub, uint32_t n_edge,
                             double seg_len, int32_t *nd, int32_t *first, int static_hit) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edge) return;
    double a[TMK_MAXSUB], b[TMK_MAXSUB];
    for (int k = 0; k < n_sub; k++) {
        a[k] = load_q<double>(p0, layout, ld, e, k);
        b[k] = load_q<double>(p1, layout, ld, e, k);
    }
    nd[e] = segment_count(a, b, n_sub, seg_len);
    first[e] = static_hit ? 0 : TMK_FIRST_NONE;
}

/* survivors of level L that have states at level L+1 (nd > 2^L) */
__global__ void k_edges_compact(const uint32_t *alive_in, const uint32_t *n_in, uint32_t n_edge, const int32_t *nd,
                                const int32_t *first, int level, uint32_t *alive_out, uint32_t *n_out) {
    const uint32_t na = alive_in ? *n_in : n_edge;
    const int lane = threadIdx.x & 31;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < na; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        bool keep = false;
        uint32_t e = 0;
        if (i < na) {
            e = alive_in ? alive_in[i] : i;
            keep = first[e] == TMK_FIRST_NONE && nd[e] > (1 << level);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            uint32_t off = 0;
            if (lane == 0) off = atomicAdd(n_out, (uint32_t)__popc(m));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (keep) alive_out[off + __popc(m & ((1u << lane) - 1u))] = e;
        }
    }
}

/* all edges sequentially in double - only when the undecided list overflowed */
__global__ void __launch_bounds__(128) k_edges_overflow_all(Image<double> im, SrcEdgeLevel src, UncList unc, int32_t *first) {
    if (*unc.overflow == 0) return;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < src.n_edge; e += gridDim.x * blockDim.x) {
        int f = TMK_FIRST_NONE;
        if (im.static_hit) f = 0;
        const int nd = src.nd[e];
        for (int pos = 0; pos < nd && f == TMK_FIRST_NONE; pos++) {
            const SrcEdgeLevel::State st = src.at(e, pos);
            double q[TMK_MAXSUB];
            for (int k = 0; k < im.n_sub; k++) q[k] = src.load1<double>(st, k);
            if (eval_config(im, q, (uint8_t *)nullptr, (DevStats *)nullptr) == HIT) f = pos;
        }
        first[e] = f;
    }
}

/* validity bitmap + first_invalid output from first[] */
__global__ void k_edges_finish(const int32_t *first, uint32_t n_edge, uint32_t *bits, int32_t *first_out) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < n_edge && first[e] == TMK_FIRST_NONE;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && e < n_edge) bits[e >> 5] = m;
    if (first_out && e < n_edge) first_out[e] = valid ? -1 : first[e];
}

static inline int align16(int x) { return (x + 15) & ~15; }

#define TMK_STAGE_BUDGET (40 * 1024) /* images up to this size are staged in shared memory */

template <class Kernel> static void warp_configure(WarpCfg &cfg, const Image<float> &im, int n_jslab, Kernel kernel) {
    cfg.bytes_joint = align16(im.n_joint * (int)sizeof(DJoint<float>));
    cfg.bytes_mg = align16(im.n_mg * (int)sizeof(DGeom<float>));
    cfg.bytes_sg = align16(im.n_sg * (int)sizeof(DGeom<float>));
    cfg.bytes_pairs = align16(im.n_pair * 4);
    cfg.bytes_sq = align16(im.n_sg * (int)sizeof(SQuick));
    cfg.bytes_jg = align16((im.n_joint + 1) * 4);
    cfg.bytes_pa = align16((im.n_mg + 1) * 4);
    int image = cfg.bytes_joint + cfg.bytes_mg + cfg.bytes_sg + cfg.bytes_pairs + cfg.bytes_sq + cfg.bytes_jg + cfg.bytes_pa;
    cfg.stage = image <= TMK_STAGE_BUDGET ? 1 : 0;
    int off = 0;
    cfg.off_joint = off; off += cfg.stage ? cfg.bytes_joint : 0;
    cfg.off_mg = off; off += cfg.stage ? cfg.bytes_mg : 0;
    cfg.off_sg = off; off += cfg.stage ? cfg.bytes_sg : 0;
    cfg.off_pairs = off; off += cfg.stage ? cfg.bytes_pairs : 0;
    cfg.off_sq = off; off += cfg.stage ? cfg.bytes_sq : 0;
    cfg.off_jg = off; off += cfg.stage ? cfg.bytes_jg : 0;
    cfg.off_pa = off; off += cfg.stage ? cfg.bytes_pa : 0;
    off = (off + 127) & ~127;
    cfg.off_warp = off;
    int w = 0;
    w += im.n_mg * 7 * 32 * 4;
    cfg.woff_q = w; w += im.n_sub * 32 * 4;
    cfg.woff_queue = w; w += TMK_WQ * 4;
    cfg.woff_flags = w; w += 32;
    cfg.woff_unc = w; w += 64 * 4;
    cfg.woff_jslab = w; w += n_jslab * 7 * 32 * 4;
    cfg.warp_bytes = (w + 127) & ~127;
    cfg.smem = cfg.off_warp + TMK_WARPS * cfg.warp_bytes;
    cfg.grid = 0;
    int dev = 0, sms = 148, per_sm = 1, optin = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cfg.smem = -1; return; }
    int max_dyn = optin - (int)fa.sharedSizeBytes;
    /* the attribute is per function, not per collision context: always raise it to the maximum */
    if (cfg.smem > max_dyn || cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) != cudaSuccess) {
        cfg.smem = -1;
        return;
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TMK_WARPS * 32, cfg.smem);
    if (per_sm < 1) per_sm = 1;
    cfg.grid = sms * per_sm;
}

} /* namespace tmk */
