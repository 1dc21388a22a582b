/* pipeline.cuh - the batched state-validity pipeline (K1 + K2 + K3 fused; K4 on top of it).
 *
 * Persistent CTAs of TILE threads walk tiles of TILE states; the phases are CTA-synchronous so that
 * the warps of an SM execute the same small code region at a time (instruction-cache capacity, not
 * the FMA pipe, is what this branchy code runs out of - see profiles/README.md):
 *
 *   stage   scene tables (joints, geometry, grouped pair list, per-pair bounding spheres, quick tables)
 *           copied once per CTA into shared memory with TMA 1-D bulk copies completing on an
 *           mbarrier (images above the staging budget are read through L1 instead)
 *   P1+P2   one thread = one state.  The joint chain runs in registers; as soon as the pose of a
 *           moving geometry exists it is written to the pose slab ([geom*7+k][thread], bank conflict
 *           free) and meets its partner list: static partners by bounding sphere, static boxes and
 *           meshes by the exact sphere-to-oriented-box distance, moving partners from the slab.  The
 *           list is warp-uniform; survivors are compacted with ballot/popc into the warp's private
 *           region of the queue (fill counts are warp-uniform registers, no atomics).
 *   Q       one thread = one surviving (state, pair): quick certificates (dev_math.cuh quick_pair).
 *   G       exact convex narrowphase (SAT / GJK) on what the certificates left open.
 *   M       mesh narrowphase: AABB-tree traversal, per-triangle certificates, GJK on the rest.
 *   out     a state is valid unless a HIT was certified.  Pairs or single mesh triangles the FP32
 *           certificates cannot decide (UNC) are appended to a global list; k_recheck_items
 *           re-evaluates exactly those in double precision and clears the state's result on a HIT.
 *
 * Src supplies the states (configurations, or interior states of swept edges); Sink consumes the
 * verdicts (validity bitmap, or per-edge first-invalid positions).
 */
#pragma once
#include "kernels.cuh"

namespace tmk {

/* world-frame quick geometry: centre, axes (columns of the rotation), dims, bounding radius */
struct SQuick {
    float cx, cy, cz, brad;
    float a0x, a0y, a0z, h0; /* axis i and the dimension along it share one 16-byte load */
    float a1x, a1y, a1z, h1;
    float a2x, a2y, a2z, h2;
};

struct TileCfg {
    int stage; /* 1: image staged in shared memory */
    int off_joint, off_mg, off_sg, off_pairs, off_sq, off_jg, off_pa, off_ps4;
    int bytes_joint, bytes_mg, bytes_sg, bytes_pairs, bytes_sq, bytes_jg, bytes_pa, bytes_ps4;
    int off_gcell, off_gitems, off_gc4, bytes_gcell, bytes_gitems, bytes_gc4; /* static grid (large scenes) */
    int off_slab, off_qs, off_jslab, off_q1, off_q2, off_q3, off_flags, off_unc;
    int q1_cap, q2_cap, q3_cap;
    int tile; /* threads per CTA = states per pass: 256 or 128 */
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
    g.x = mk<float>(s.a0x, s.a0y, s.a0z);
    g.y = mk<float>(s.a1x, s.a1y, s.a1z);
    g.z = mk<float>(s.a2x, s.a2y, s.a2z);
    g.h0 = s.h0; g.h1 = s.h1; g.h2 = s.h2;
    g.brad = s.brad;
    return g;
}

/* queue item: bits 23-30 thread of the tile | 17-22 moving geometry a | 16 partner is static | 0-15 partner b.
 * The low 23 bits ("pair key") also identify the pair in the undecided list. */
__device__ __forceinline__ uint32_t item_pack(int tid, int a, int b, int b_static) {
    return ((uint32_t)tid << 23) | ((uint32_t)a << 17) | ((uint32_t)b_static << 16) | (uint32_t)b;
}
#define ITEM_T(e) ((int)((e) >> 23))
#define ITEM_A(e) ((int)(((e) >> 17) & 63))
#define ITEM_STATIC(e) ((int)(((e) >> 16) & 1))
#define ITEM_B(e) ((int)((e)&0xffff))
#define ITEM_KEY(e) ((e)&0x7fffffu)

struct UncList {
    uint4 *items;       /* x,y = state identity (Src::unc_id), z = pair key (ITEM_KEY), w = mesh triangle or TMK_WHOLE_PAIR */
    uint32_t *n;        /* appended count (may exceed cap) */
    uint32_t cap;
    uint32_t *overflow; /* set when an item was dropped: the brute-force double kernel then runs */
    uint32_t *spill;    /* per-CTA overflow area of the broadphase queue in global memory: [gridDim.x][spill_cap] */
    uint32_t spill_cap;
    /* deferred mesh narrowphase: the pipeline exports its surviving mesh items (with the poses they need)
     * and k_heavy_items runs them with far more items in flight per SM than a phase of the tile loop can */
    float4 *mesh_items; /* TMK_MESH_REC float4 per item: mesh pairs */
    uint32_t *mesh_n;
    uint32_t mesh_cap;
    float4 *conv_items; /* same records: convex pairs for SAT / GJK */
    uint32_t *conv_n;
    uint32_t conv_cap;
    uint2 *tri_items;   /* (mesh record, triangle) the per-triangle certificates left open */
    uint32_t *tri_n;
    uint32_t tri_cap;
    uint32_t *tile_ctr; /* [0] next tile of the running k_states_tile launch, [1] CTAs that finished it */
};
#define TMK_MESH_REC 5 /* float4 per exported mesh item: (id_x, id_y, key, -) poseA[7] - poseB[7] - */
#define TMK_WHOLE_PAIR 0xffffffffu
__device__ __forceinline__ void unc_emit(const UncList &u, uint32_t id_x, uint32_t id_y, uint32_t pair, uint32_t tri) {
    uint32_t i = atomicAdd(u.n, 1u);
    if (i < u.cap) u.items[i] = make_uint4(id_x, id_y, pair, tri);
    else *u.overflow = 1u;
}
/* mesh narrowphase hands its undecided TRIANGLES over one by one (the exact re-check of a whole
 * mesh pair in one thread is a long serial chain; single triangles are not) */
struct TriNeed { /* k_mesh_traverse: open triangles go to the triangle list */
    static constexpr bool enabled = true;
    UncList u;
    uint32_t rec;
    uint32_t id_x, id_y, pair;
    __device__ __forceinline__ void operator()(int tri) const {
        const uint32_t i = atomicAdd(u.tri_n, 1u);
        if (i < u.tri_cap) u.tri_items[i] = make_uint2(rec, (uint32_t)tri);
        else unc_emit(u, id_x, id_y, pair, (uint32_t)tri); /* list full: decide the triangle exactly */
    }
};
struct TriEmit {
    static constexpr bool enabled = true;
    UncList u;
    uint32_t id_x, id_y, pair;
    __device__ __forceinline__ void operator()(int tri) const { unc_emit(u, id_x, id_y, pair, (uint32_t)tri); }
};

/* the rare path of the broadphase queues, out of line so that the hot loops stay small: a full
 * shared-memory region spills into the CTA's global area (same float phases later); only when that is
 * full too - or for a mesh item - is the pair handed to the exact re-check */
__device__ __noinline__ void queue_overflow(const UncList unc, uint32_t *spill_n, uint32_t *spill, const uint32_t *unc_x,
                                            const uint32_t *unc_y, uint32_t item, bool mesh) {
    if (!mesh) {
        const uint32_t k = atomicAdd(spill_n, 1u);
        if (k < unc.spill_cap) { spill[k] = item; return; }
    }
    unc_emit(unc, unc_x[ITEM_T(item)], unc_y[ITEM_T(item)], ITEM_KEY(item), TMK_WHOLE_PAIR);
}

/* ------------------------------------------------------------------ sources of states, sinks of verdicts
 * Src:  count(), prepare(s) -> State (with .live), load1<T>(State, k) = k-th joint value,
 *       unc_id(State) / from_id(x, y) to carry a state's identity through the undecided list.
 * Sink: emit(State, g, hit, lane) called convergently by the 32 lanes of group g; mark_hit(State). */
struct SrcConfigs {
    const void *p;
    int layout;
    size_t ld, n;
    size_t s_begin; /* this launch covers states s_begin .. s_begin + n - 1 of the batch at p */
    struct State { size_t s; bool live; };
    __device__ __forceinline__ size_t count() const { return n; }
    __device__ __forceinline__ State prepare(size_t s) const { State st; st.s = s_begin + s; st.live = s < n; return st; }
    template <typename T> __device__ __forceinline__ T load1(const State &st, int k) const {
        return load_q<T>(p, layout, ld, st.s, k);
    }
    __device__ __forceinline__ uint2 unc_id(const State &st) const { return make_uint2((uint32_t)st.s, 0u); }
    __device__ __forceinline__ State from_id(uint32_t x, uint32_t y) const { (void)y; State st; st.s = x; st.live = true; return st; }
};

struct SinkBits {
    uint32_t *bits;
    __device__ __forceinline__ void emit(const SrcConfigs::State &st, size_t g, bool hit, int lane) const {
        (void)g;
        const unsigned valid = __ballot_sync(0xffffffffu, !hit);
        /* lane 0 holds the warp's first state: if it lies past the batch the whole warp does, and its word is
         * past ceil(n / 32) - the caller's bitmap has no padding (slices of a sharded bitmap touch each other) */
        if (lane == 0 && st.live) bits[st.s >> 5] = valid;
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
    __device__ __forceinline__ uint2 unc_id(const State &st) const { return make_uint2(st.e, (uint32_t)st.pos); }
    __device__ __forceinline__ State from_id(uint32_t x, uint32_t y) const { return at(x, (int)y); }
};

/* Small edge batches (a planner's frontier: a hundred edges of ten states): ALL states of all edges in one launch.
 * Level by level the batch is a chain of ~20 dependent launches of a few microseconds of work each; without the
 * early-out of the levels it is a few thousand states - one tile launch.  State s belongs to the edge e with
 * off[e] <= s < off[e+1] (off = exclusive prefix sum of the segment counts, built by k_edges_prep_flat) at position
 * s - off[e] of OMPL's test order; first[e] = atomicMin of the invalid positions, exactly as for the levels. */
struct SrcEdgeFlat {
    SrcEdgeLevel lv;       /* end points, layout, segment counts: at(), load1(), ids as for the levels */
    const uint32_t *off;   /* [n_edge + 1] */
    typedef SrcEdgeLevel::State State;
    __device__ __forceinline__ size_t count() const { return (size_t)off[lv.n_edge]; }
    __device__ __forceinline__ State prepare(size_t s) const {
        const uint32_t n = lv.n_edge;
        if (s >= (size_t)off[n]) { State st; st.e = 0; st.pos = 0; st.nd = 0; st.t = 1.0; st.live = false; return st; }
        uint32_t lo = 0, hi = n; /* largest e with off[e] <= s */
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((size_t)off[mid] <= s) lo = mid; else hi = mid;
        }
        return lv.at(lo, (int)(s - off[lo]));
    }
    template <typename T> __device__ __forceinline__ T load1(const State &st, int k) const { return lv.template load1<T>(st, k); }
    __device__ __forceinline__ uint2 unc_id(const State &st) const { return lv.unc_id(st); }
    __device__ __forceinline__ State from_id(uint32_t x, uint32_t y) const { return lv.from_id(x, y); }
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

#ifdef TMK_PHASE_CLOCKS
/* tuning aid (not in the product build): cycles per phase, summed over CTAs by thread 0 */
__device__ unsigned long long g_phase_cycles[8];
#define PHASE_MARK(k)                                                   \
    do {                                                                \
        if (tid == 0) {                                                 \
            const long long now_ = clock64();                           \
            atomicAdd(&g_phase_cycles[k], (unsigned long long)(now_ - t_phase)); \
            t_phase = now_;                                             \
        }                                                               \
    } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

/* ------------------------------------------------------------------ the kernel
 * TILE threads = TILE states per CTA pass; phases are CTA-synchronous so that all warps of an SM
 * execute the same (small) code region at a time - the instruction caches (L0 6 KB, L1.5 32 KB)
 * are the scarce resource of this branchy code, not the FMA pipe. */
template <class Src, class Sink, int TILE, bool GRID>
#ifndef TMK_TILE_MINCTAS
#define TMK_TILE_MINCTAS(T) (512 / (T))
#endif
__global__ void __launch_bounds__(TILE, TMK_TILE_MINCTAS(TILE))
k_states_tile(Image<float> im, TileCfg cfg, Src src, Sink sink, UncList unc, unsigned long long *n_states) {
    const size_t n = src.count();
    if (n == 0) return; /* dead bisection levels cost a launch, not a staging */
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t qn[4];          /* [1]: fill of q2 (convex narrowphase), [3]: fill of the global spill area */
    __shared__ uint32_t wn1[TILE / 32]; /* fill of each warp's private region of q1 (quick) ... */
    __shared__ uint32_t wn3[TILE / 32]; /* ... and of q3 (mesh narrowphase) */

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DJoint<float> *J = im.joints;
    const DGeom<float> *MG = im.mg;
    const DGeom<float> *SG = im.sg;
    const uint32_t *PAIRS = im.t_pairs; /* the pipeline's list (without the pairs the static grid covers) */
    const SQuick *SQ = (const SQuick *)im.sq;
    const int *JG = im.jg_begin, *PA = im.pa_begin;
    const float4 *PS4 = (const float4 *)im.ps4;
    const int *GCELL = im.grid.cell_start;
    const unsigned short *GITEMS = im.grid.items;
    const float4 *GC4 = im.grid.c4;

    if (cfg.stage) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t total = (uint32_t)(cfg.bytes_joint + cfg.bytes_mg + cfg.bytes_sg + cfg.bytes_pairs + cfg.bytes_sq +
                                        cfg.bytes_jg + cfg.bytes_pa + cfg.bytes_ps4);
            if (GRID) total += (uint32_t)(cfg.bytes_gcell + cfg.bytes_gitems + cfg.bytes_gc4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(total)
                         : "memory");
            bulk_g2s(smem + cfg.off_joint, im.joints, (uint32_t)cfg.bytes_joint, &bar);
            bulk_g2s(smem + cfg.off_mg, im.mg, (uint32_t)cfg.bytes_mg, &bar);
            if (cfg.bytes_sg) bulk_g2s(smem + cfg.off_sg, im.sg, (uint32_t)cfg.bytes_sg, &bar);
            if (cfg.bytes_pairs) bulk_g2s(smem + cfg.off_pairs, im.t_pairs, (uint32_t)cfg.bytes_pairs, &bar);
            if (cfg.bytes_sq) bulk_g2s(smem + cfg.off_sq, im.sq, (uint32_t)cfg.bytes_sq, &bar);
            bulk_g2s(smem + cfg.off_jg, im.jg_begin, (uint32_t)cfg.bytes_jg, &bar);
            bulk_g2s(smem + cfg.off_pa, im.pa_begin, (uint32_t)cfg.bytes_pa, &bar);
            if (cfg.bytes_ps4) bulk_g2s(smem + cfg.off_ps4, im.ps4, (uint32_t)cfg.bytes_ps4, &bar);
            if (GRID && cfg.bytes_gcell) {
                bulk_g2s(smem + cfg.off_gcell, im.grid.cell_start, (uint32_t)cfg.bytes_gcell, &bar);
                bulk_g2s(smem + cfg.off_gitems, im.grid.items, (uint32_t)cfg.bytes_gitems, &bar);
                bulk_g2s(smem + cfg.off_gc4, im.grid.c4, (uint32_t)cfg.bytes_gc4, &bar);
            }
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
        if (cfg.bytes_sg) SG = (const DGeom<float> *)(smem + cfg.off_sg);
        if (cfg.bytes_pairs) PAIRS = (const uint32_t *)(smem + cfg.off_pairs);
        if (cfg.bytes_sq) SQ = (const SQuick *)(smem + cfg.off_sq);
        JG = (const int *)(smem + cfg.off_jg);
        PA = (const int *)(smem + cfg.off_pa);
        if (cfg.bytes_ps4) PS4 = (const float4 *)(smem + cfg.off_ps4);
        if (GRID && cfg.bytes_gcell) {
            GCELL = (const int *)(smem + cfg.off_gcell);
            GITEMS = (const unsigned short *)(smem + cfg.off_gitems);
            GC4 = (const float4 *)(smem + cfg.off_gc4);
        }
    }

    float *slab = (float *)(smem + cfg.off_slab);       /* [n_mg*7][TILE] geometry poses */
    float *qs = (float *)(smem + cfg.off_qs);           /* [n_sub][TILE] joint values */
    float *jslab = (float *)(smem + cfg.off_jslab);     /* poses of joints with non-consecutive children */
    uint32_t *q1 = (uint32_t *)(smem + cfg.off_q1);
    uint32_t *q2 = (uint32_t *)(smem + cfg.off_q2);
    uint32_t *q3 = (uint32_t *)(smem + cfg.off_q3);
    volatile uint8_t *f_hit = smem + cfg.off_flags;     /* [TILE] */
    uint32_t *unc_x = (uint32_t *)(smem + cfg.off_unc); /* [TILE] + [TILE]: undecided-item words of each state */
    uint32_t *unc_y = unc_x + TILE;

    const size_t n_tiles = (n + TILE - 1) / TILE;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_joint = im.n_joint, n_sub = im.n_sub;
    const float pad = Num<float>::eps() * 4.0f;
    constexpr bool grid_on = GRID; /* compile-time: scenes without a static grid pay nothing for it */

    uint32_t *spill = unc.spill + (size_t)blockIdx.x * unc.spill_cap;
    auto spill_or_exact = [&](uint32_t item, bool mesh) { queue_overflow(unc, &qn[3], spill, unc_x, unc_y, item, mesh); };

    auto pose_at = [&](int g, int t) {
        const float *o = slab + (size_t)g * 7 * TILE + t;
        Xf<float> X;
        X.r.x = o[0]; X.r.y = o[TILE]; X.r.z = o[2 * TILE]; X.r.w = o[3 * TILE];
        X.v.x = o[4 * TILE]; X.v.y = o[5 * TILE]; X.v.z = o[6 * TILE];
        return X;
    };

#ifdef TMK_PHASE_CLOCKS
    long long t_phase = clock64();
#endif
    /* tiles are handed out through a counter, not by CTA index: a CTA that becomes resident late (another
     * kernel - a collective on a second stream - still holds its SM) then simply takes fewer tiles instead
     * of stretching the launch by its delay */
    __shared__ uint32_t s_tile;
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(unc.tile_ctr, 1u);
        __syncthreads();
        const size_t tile = s_tile;
        if (tile >= n_tiles) break;
        const size_t s = tile * TILE + tid;
        const typename Src::State st = src.prepare(s);
        const bool live = st.live && !im.static_hit;
        if (tid < 4) qn[tid] = 0;
        if (lane == 0) { wn1[warp] = 0; wn3[warp] = 0; }
        f_hit[tid] = (st.live && im.static_hit) ? 1 : 0;
        {
            const uint2 it = src.unc_id(st);
            unc_x[tid] = it.x; unc_y[tid] = it.y;
        }
        if (st.live)
            for (int k = 0; k < n_sub; k++) qs[k * TILE + tid] = src.template load1<float>(st, k);
        if (n_states) {
            const unsigned lm = __ballot_sync(0xffffffffu, live);
            if (lane == 0 && lm) atomicAdd(n_states, (unsigned long long)__popc(lm));
        }
        __syncthreads();
        PHASE_MARK(0);

        /* ---- P1 + P2: joint chain in registers; every moving geometry meets its partner list at once ---- */
        {
            uint32_t n1 = 0, n3 = 0; /* warp-uniform */
            uint32_t *q1w = q1 + (size_t)warp * cfg.q1_cap, *q3w = q3 + (size_t)warp * cfg.q3_cap;
            Xf<float> cur;
            cur.r.x = cur.r.y = cur.r.z = 0.f; cur.r.w = 1.f; cur.v = mk<float>(0.f, 0.f, 0.f);
            for (int j = 0; j < n_joint; j++) {
                const DJoint<float> &Jt = J[j];
                const int par_id = Jt.parent;
                Xf<float> par = cur;
                if (par_id >= 0 && par_id != j - 1) {
                    const float *o = jslab + (size_t)(J[par_id].pad - 1) * 7 * TILE + tid;
                    par.r.x = o[0]; par.r.y = o[TILE]; par.r.z = o[2 * TILE]; par.r.w = o[3 * TILE];
                    par.v.x = o[4 * TILE]; par.v.y = o[5 * TILE]; par.v.z = o[6 * TILE];
                }
                cur = fk_joint(xload(Jt.E), mk<float>(Jt.ax[0], Jt.ax[1], Jt.ax[2]), Jt.off, Jt.type, qs[Jt.qslot * TILE + tid],
                               par_id >= 0 ? &par : (const Xf<float> *)nullptr);
                if (Jt.pad > 0) {
                    float *o = jslab + (size_t)(Jt.pad - 1) * 7 * TILE + tid;
                    o[0] = cur.r.x; o[TILE] = cur.r.y; o[2 * TILE] = cur.r.z; o[3 * TILE] = cur.r.w;
                    o[4 * TILE] = cur.v.x; o[5 * TILE] = cur.v.y; o[6 * TILE] = cur.v.z;
                }
                for (int a = JG[j]; a < JG[j + 1]; a++) {
                    const DGeom<float> &A = MG[a];
                    const Xf<float> XA = xmul(cur, xload(A.tf));
                    {
                        float *o = slab + (size_t)a * 7 * TILE + tid;
                        o[0] = XA.r.x; o[TILE] = XA.r.y; o[2 * TILE] = XA.r.z; o[3 * TILE] = XA.r.w;
                        o[4 * TILE] = XA.v.x; o[5 * TILE] = XA.v.y; o[6 * TILE] = XA.v.z;
                    }
                    const Vec3<float> ca = bound_centre(A, XA);
                    const float ra = A.brad + pad;
                    /* survivors of the warp-uniform tests go to the warp's own region of q1 (convex pairs)
                     * or q3 (mesh pairs): the fill counts are warp-uniform registers, no atomics */
                    auto enqueue = [&](int p, bool pass) {
                        const unsigned m = __ballot_sync(0xffffffffu, pass);
                        if (m) {
                            const uint32_t w = PAIRS[p];
                            const bool mesh = (w >> 28) == PC_MESH;
                            uint32_t base = mesh ? n3 : n1;
                            if (grid_on) { /* the divergent grid walk also appends: the fill counts live in shared memory */
                                if (lane == 0) base = atomicAdd(mesh ? &wn3[warp] : &wn1[warp], (uint32_t)__popc(m));
                                base = __shfl_sync(0xffffffffu, base, 0);
                            }
                            const uint32_t idx = base + __popc(m & lt_mask);
                            if (pass) {
                                const uint32_t item = item_pack(tid, a, (int)((w >> 8) & 0xffff), (int)((w >> 24) & 1));
                                if (!mesh && idx < (uint32_t)cfg.q1_cap) q1w[idx] = item;
                                else if (mesh && idx < (uint32_t)cfg.q3_cap) q3w[idx] = item;
                                else spill_or_exact(item, mesh);
                            }
                            if (mesh) n3 += __popc(m); else n1 += __popc(m);
                        }
                    };
                    int p = PA[3 * a];
                    const int e0 = PA[3 * a + 1], e1 = PA[3 * a + 2], e2 = PA[3 * a + 3];
                    /* static partners by bounding sphere */
                    for (; p + 4 <= e0; p += 4) { /* four tests, one vote: most groups have no survivor */
                        bool pass[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float4 c4 = PS4[p + u];
                            const float dx = c4.x - ca.x, dy = c4.y - ca.y, dz = c4.z - ca.z;
                            pass[u] = live && dx * dx + dy * dy + dz * dz <= c4.w; /* .w = (ra + pad + rb)^2 */
                        }
                        if (__any_sync(0xffffffffu, pass[0] | pass[1] | pass[2] | pass[3])) {
#pragma unroll
                            for (int u = 0; u < 4; u++) enqueue(p + u, pass[u]);
                        }
                    }
                    for (; p < e0; p++) {
                        const float4 c4 = PS4[p];
                        const float dx = c4.x - ca.x, dy = c4.y - ca.y, dz = c4.z - ca.z;
                        enqueue(p, live && dx * dx + dy * dy + dz * dz <= c4.w);
                    }
                    /* static boxes: bounding sphere of A against the box itself (exact point-box distance) */
                    auto box_pass = [&](int pp) {
                        const int b = (PAIRS[pp] >> 8) & 0xffff;
                        const float4 c4 = *reinterpret_cast<const float4 *>(&SQ[b].cx);
                        const float4 x4 = *reinterpret_cast<const float4 *>(&SQ[b].a0x);
                        const float4 y4 = *reinterpret_cast<const float4 *>(&SQ[b].a1x);
                        const float4 z4 = *reinterpret_cast<const float4 *>(&SQ[b].a2x);
                        const float dx = c4.x - ca.x, dy = c4.y - ca.y, dz = c4.z - ca.z;
                        const float ex = fmaxf(fabsf(dx * x4.x + dy * x4.y + dz * x4.z) - x4.w, 0.f);
                        const float ey = fmaxf(fabsf(dx * y4.x + dy * y4.y + dz * y4.z) - y4.w, 0.f);
                        const float ez = fmaxf(fabsf(dx * z4.x + dy * z4.y + dz * z4.z) - z4.w, 0.f);
                        return live && ex * ex + ey * ey + ez * ez <= ra * ra;
                    };
                    /* boxes aligned with the world axes come first in the segment: centre + half extents (Image::ab4,
                     * read through L1: a warp-uniform 16-byte load), half the arithmetic of the oriented test */
                    {
                        const float4 *AB4 = (const float4 *)im.ab4;
                        const float ra2 = ra * ra;
                        auto aabb_pass = [&](int pp) {
                            const float4 c4 = AB4[2 * pp], h4 = AB4[2 * pp + 1];
                            const float ex = fmaxf(fabsf(c4.x - ca.x) - h4.x, 0.f), ey = fmaxf(fabsf(c4.y - ca.y) - h4.y, 0.f),
                                        ez = fmaxf(fabsf(c4.z - ca.z) - h4.z, 0.f);
                            return live && ex * ex + ey * ey + ez * ez <= ra2;
                        };
                        const int e0a = im.pa_aabb[a];
                        for (; p + 4 <= e0a; p += 4) {
                            bool pass[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) pass[u] = aabb_pass(p + u);
                            if (__any_sync(0xffffffffu, pass[0] | pass[1] | pass[2] | pass[3])) {
#pragma unroll
                                for (int u = 0; u < 4; u++) enqueue(p + u, pass[u]);
                            }
                        }
                        for (; p < e0a; p++) enqueue(p, aabb_pass(p));
                    }
                    for (; p + 2 <= e1; p += 2) {
                        const bool p0 = box_pass(p), p1 = box_pass(p + 1);
                        if (__any_sync(0xffffffffu, p0 | p1)) { enqueue(p, p0); enqueue(p + 1, p1); }
                    }
                    for (; p < e1; p++) enqueue(p, box_pass(p));
                    /* moving partners (their poses are already in the slab) */
                    for (; p < e2; p++) {
                        const int b = (PAIRS[p] >> 8) & 0xffff;
                        const DGeom<float> &B = MG[b];
                        const float *o = slab + (size_t)b * 7 * TILE + tid;
                        Vec3<float> cb = mk<float>(o[4 * TILE], o[5 * TILE], o[6 * TILE]);
                        if (B.shape == K_MESH) cb = bound_centre(B, pose_at(b, tid));
                        const Vec3<float> d = cb - ca;
                        const float reach = ra + B.brad;
                        enqueue(p, live && dot(d, d) <= reach * reach);
                    }
                    /* small static geometry of large scenes: walk the grid cells the bounding sphere can reach
                     * (lane-divergent: every lane has its own cells) */
                    if (grid_on && live) {
                        const StaticGrid &G = im.grid;
                        const float rq = ra + G.r_max;
                        const int x0 = max(0, (int)floorf((ca.x - rq - G.ox) * G.inv_h)), x1 = min(G.nx - 1, (int)floorf((ca.x + rq - G.ox) * G.inv_h));
                        const int y0 = max(0, (int)floorf((ca.y - rq - G.oy) * G.inv_h)), y1 = min(G.ny - 1, (int)floorf((ca.y + rq - G.oy) * G.inv_h));
                        const int z0 = max(0, (int)floorf((ca.z - rq - G.oz) * G.inv_h)), z1 = min(G.nz - 1, (int)floorf((ca.z + rq - G.oz) * G.inv_h));
                        const uint32_t *live_row = G.live + (size_t)a * G.wps;
                        for (int z = z0; z <= z1; z++)
                            for (int y = y0; y <= y1; y++) {
                                const int row = (z * G.ny + y) * G.nx;
                                const int k1 = GCELL[row + x1 + 1];
                                for (int k = GCELL[row + x0]; k < k1; k++) { /* cells x0..x1 of a row are contiguous */
                                    const float4 c4 = GC4[k];
                                    const float dx = c4.x - ca.x, dy = c4.y - ca.y, dz = c4.z - ca.z;
                                    const float reach = ra + c4.w;
                                    if (dx * dx + dy * dy + dz * dz > reach * reach) continue;
                                    const int b = GITEMS[k]; /* rare from here on */
                                    if (!((live_row[b >> 5] >> (b & 31)) & 1u)) continue;
                                    const uint32_t idx = atomicAdd(&wn1[warp], 1u);
                                    const uint32_t item = item_pack(tid, a, b, 1);
                                    if (idx < (uint32_t)cfg.q1_cap) q1w[idx] = item;
                                    else spill_or_exact(item, false);
                                }
                            }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                wn1[warp] = min(grid_on ? wn1[warp] : n1, (uint32_t)cfg.q1_cap);
                wn3[warp] = min(grid_on ? wn3[warp] : n3, (uint32_t)cfg.q3_cap);
            }
        }
        __syncthreads();
        PHASE_MARK(1);

        /* ---- Q: quick certificates, one thread per surviving (state, pair) ---- */
        {
            auto quick_item = [&](const uint32_t e) {
                const int t = ITEM_T(e), a = ITEM_A(e), b = ITEM_B(e);
                if (f_hit[t]) return;
                const DGeom<float> &A = MG[a];
                const QG qa = qg_from_pose(A, pose_at(a, t));
                int r;
                if (ITEM_STATIC(e)) r = quick_pair(A.shape, qa, SG[b].shape, qg_from_static(SQ[b]));
                else r = quick_pair(A.shape, qa, MG[b].shape, qg_from_pose(MG[b], pose_at(b, t)));
                if (r == HIT) f_hit[t] = 1;
                else if (r == UNC) unc_emit(unc, unc_x[t], unc_y[t], ITEM_KEY(e), TMK_WHOLE_PAIR);
                else if (r == NEED) {
                    const uint32_t idx = atomicAdd(&qn[1], 1u);
                    if (idx < (uint32_t)cfg.q2_cap) q2[idx] = e;
                    else unc_emit(unc, unc_x[t], unc_y[t], ITEM_KEY(e), TMK_WHOLE_PAIR);
                }
            };
            uint32_t total = 0;
#pragma unroll
            for (int k = 0; k < TILE / 32; k++) total += wn1[k];
            const uint32_t n_spill = min(qn[3], unc.spill_cap);
            for (uint32_t i = tid; i < total + n_spill; i += TILE) {
                uint32_t e;
                if (i < total) { /* dense index -> (warp region, offset) */
                    uint32_t off = i;
                    int k = 0;
                    while (off >= wn1[k]) { off -= wn1[k]; k++; }
                    e = q1[(size_t)k * cfg.q1_cap + off];
                } else e = spill[i - total];
                quick_item(e);
            }
        }
        __syncthreads();
        PHASE_MARK(2);

        /* ---- X: export the heavy items - pairs the certificates left open (SAT / GJK) and mesh pairs (BVH
         * traversal) - of states that are still undecided.  As phases of this loop they were 60 % of the
         * tile time at 25 % of the instructions (few items, long dependent chains, everybody else waiting
         * at the barrier); the wavefront kernels below run them with thousands of items in flight per SM ---- */
        {
            uint32_t total3 = 0;
#pragma unroll
            for (int k = 0; k < TILE / 32; k++) total3 += wn3[k];
            const uint32_t total2 = min(qn[1], (uint32_t)cfg.q2_cap);
            const uint32_t total = total2 + total3;
            for (uint32_t base = 0; base < total; base += TILE) { /* uniform trip count: the ballots are convergent */
                const uint32_t i = base + tid;
                bool keep = false;
                uint32_t e = 0;
                const bool conv = i < total2;
                if (conv) e = q2[i];
                else if (i < total) {
                    uint32_t off = i - total2;
                    int k = 0;
                    while (off >= wn3[k]) { off -= wn3[k]; k++; }
                    e = q3[(size_t)k * cfg.q3_cap + off];
                }
                if (i < total) keep = f_hit[ITEM_T(e)] == 0;
                const unsigned mc = __ballot_sync(0xffffffffu, keep && conv), mm = __ballot_sync(0xffffffffu, keep && !conv);
                uint32_t dc = 0, dm = 0;
                if (lane == 0 && mc) dc = atomicAdd(unc.conv_n, (uint32_t)__popc(mc));
                if (lane == 0 && mm) dm = atomicAdd(unc.mesh_n, (uint32_t)__popc(mm));
                dc = __shfl_sync(0xffffffffu, dc, 0) + __popc(mc & lt_mask);
                dm = __shfl_sync(0xffffffffu, dm, 0) + __popc(mm & lt_mask);
                if (keep) {
                    const int t = ITEM_T(e);
                    const uint32_t dst = conv ? dc : dm;
                    if (dst < (conv ? unc.conv_cap : unc.mesh_cap)) {
                        const Xf<float> XA = pose_at(ITEM_A(e), t);
                        Xf<float> XB = XA;
                        if (!ITEM_STATIC(e)) XB = pose_at(ITEM_B(e), t);
                        float4 *rec = (conv ? unc.conv_items : unc.mesh_items) + (size_t)dst * TMK_MESH_REC;
                        rec[0] = make_float4(__uint_as_float(unc_x[t]), __uint_as_float(unc_y[t]), __uint_as_float(ITEM_KEY(e)), 0.f);
                        rec[1] = make_float4(XA.r.x, XA.r.y, XA.r.z, XA.r.w);
                        rec[2] = make_float4(XA.v.x, XA.v.y, XA.v.z, XB.v.x);
                        rec[3] = make_float4(XB.r.x, XB.r.y, XB.r.z, XB.r.w);
                        rec[4] = make_float4(XB.v.y, XB.v.z, 0.f, 0.f);
                    } else unc_emit(unc, unc_x[t], unc_y[t], ITEM_KEY(e), TMK_WHOLE_PAIR); /* list full: decide it exactly */
                }
            }
        }
        __syncthreads();
        PHASE_MARK(4);

        sink.emit(st, tile * (TILE / 32) + warp, f_hit[tid] != 0 || !st.live, lane);
        __syncthreads();
        PHASE_MARK(5);
    }
    /* the last CTA out re-arms the counter for the next launch of the batch (launches are stream-ordered) */
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(unc.tile_ctr + 1, 1u) == gridDim.x - 1) {
            unc.tile_ctr[0] = 0;
            unc.tile_ctr[1] = 0;
            __threadfence();
        }
    }
}

/* ------------------------------------------------------------------ deferred narrowphase, wavefront style:
 * every kernel runs one kind of work with one item per thread, so that warps stay converged
 * (as one mixed kernel only 3 of 32 lanes were active per instruction) */
struct HeavyRec {
    uint32_t id_x, id_y, key;
    Xf<float> XA, XB;
};
__device__ __forceinline__ HeavyRec heavy_load(const float4 *rec) {
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4];
    HeavyRec h;
    h.id_x = __float_as_uint(r0.x); h.id_y = __float_as_uint(r0.y); h.key = __float_as_uint(r0.z);
    h.XA.r.x = r1.x; h.XA.r.y = r1.y; h.XA.r.z = r1.z; h.XA.r.w = r1.w;
    h.XA.v.x = r2.x; h.XA.v.y = r2.y; h.XA.v.z = r2.z;
    h.XB.r.x = r3.x; h.XB.r.y = r3.y; h.XB.r.z = r3.z; h.XB.r.w = r3.w;
    h.XB.v.x = r2.w; h.XB.v.y = r4.x; h.XB.v.z = r4.y;
    return h;
}

/* convex pairs the quick certificates left open: SAT / GJK */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_conv_items(Image<float> im, Src src, Sink sink, UncList unc) {
    uint32_t n = *unc.conv_n;
    if (n > unc.conv_cap) n = unc.conv_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const HeavyRec h = heavy_load(unc.conv_items + (size_t)i * TMK_MESH_REC);
        const DGeom<float> &A = im.mg[ITEM_A(h.key)];
        const DGeom<float> &B = ITEM_STATIC(h.key) ? im.sg[ITEM_B(h.key)] : im.mg[ITEM_B(h.key)];
        const Xf<float> XB = ITEM_STATIC(h.key) ? xload(B.tf) : h.XB;
        const int r = test_convex(shape_of(A), h.XA, shape_of(B), XB);
        if (r == HIT) sink.mark_hit(src.from_id(h.id_x, h.id_y));
        else if (r == UNC) unc_emit(unc, h.id_x, h.id_y, h.key, TMK_WHOLE_PAIR);
    }
}

/* mesh pairs: AABB-tree traversal with the per-triangle certificates; open triangles go to the
 * triangle list */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_mesh_traverse_simple(Image<float> im, Src src, Sink sink, UncList unc) {
    uint32_t n = *unc.mesh_n;
    if (n > unc.mesh_cap) n = unc.mesh_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const HeavyRec h = heavy_load(unc.mesh_items + (size_t)i * TMK_MESH_REC);
        const DGeom<float> &A = im.mg[ITEM_A(h.key)];
        const DGeom<float> &B = ITEM_STATIC(h.key) ? im.sg[ITEM_B(h.key)] : im.mg[ITEM_B(h.key)];
        const Xf<float> XB = ITEM_STATIC(h.key) ? xload(B.tf) : h.XB;
        TriEmit on_unc;
        on_unc.u = unc; on_unc.id_x = h.id_x; on_unc.id_y = h.id_y; on_unc.pair = h.key;
        TriNeed on_need;
        on_need.u = unc; on_need.rec = i; on_need.id_x = h.id_x; on_need.id_y = h.id_y; on_need.pair = h.key;
        const int r = test_geoms(A, h.XA, B, XB, im.meshes, on_unc, on_need);
        if (r == HIT) sink.mark_hit(src.from_id(h.id_x, h.id_y));
        else if (r == UNC) unc_emit(unc, h.id_x, h.id_y, h.key, TMK_WHOLE_PAIR);
    }
}

/* The same walk with the lanes kept busy.  One item per thread leaves 5.5 of 32 lanes active per instruction:
 * most walks end at the root box, a few visit dozens of nodes and test triangles, and a lane in a leaf
 * (vertex transforms + certificates, ~150 instructions) stalls the lanes that only step through inner nodes
 * (~30).  Here a warp is a pool of 32 walkers over a shared work counter: idle lanes refill together once
 * a quarter of the warp is idle, inner-node steps run for every lane that has no leaf pending, and pending
 * leaves are processed one triangle at a time when enough lanes hold one (or nobody can step). */
#define TMK_TRAV_STACK 24
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_mesh_traverse(Image<float> im, Src src, Sink sink, UncList unc, uint32_t *next) {
    __shared__ int stk[TMK_TRAV_STACK][128];
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t n = *unc.mesh_n;
    if (n > unc.mesh_cap) n = unc.mesh_cap;
    bool busy = false, exhausted = false, unc_whole = false;
    uint32_t rec = 0, id_x = 0, id_y = 0, key = 0;
    const BvhNode *nodes = nullptr;
    const float *tris = nullptr;
    Shape<float> S;
    S.kind = K_SPHERE; S.a = S.b = S.c = 0.f;
    MeshQuery<float> Q;
    int sp = 0, leaf_a = 0, leaf_n = 0;
    for (;;) {
        unsigned busy_m = __ballot_sync(0xffffffffu, busy);
        if (!exhausted && __popc(~busy_m) >= 8) {
            const unsigned idle = ~busy_m;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (uint32_t)__popc(idle) >= n) exhausted = true;
            if (!busy) {
                const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                if (i < n) {
                    const HeavyRec h = heavy_load(unc.mesh_items + (size_t)i * TMK_MESH_REC);
                    rec = i; id_x = h.id_x; id_y = h.id_y; key = h.key;
                    const DGeom<float> &A = im.mg[ITEM_A(key)];
                    const DGeom<float> &B = ITEM_STATIC(key) ? im.sg[ITEM_B(key)] : im.mg[ITEM_B(key)];
                    const Xf<float> XB = ITEM_STATIC(key) ? xload(B.tf) : h.XB;
                    if (A.shape == K_MESH && B.shape == K_MESH) {
                        unc_emit(unc, id_x, id_y, key, TMK_WHOLE_PAIR); /* mesh against mesh: exact path only */
                    } else {
                        const bool a_mesh = A.shape == K_MESH;
                        const MeshRef<float> &M = im.meshes[a_mesh ? A.mesh : B.mesh];
                        nodes = M.nodes; tris = M.tris;
                        S = shape_of(a_mesh ? B : A);
                        Q.init(a_mesh ? h.XA : XB, S, a_mesh ? XB : h.XA, a_mesh ? B.brad : A.brad);
                        stk[0][tid] = 0;
                        sp = 1; leaf_n = 0; unc_whole = false;
                        busy = true;
                    }
                }
            }
            busy_m = __ballot_sync(0xffffffffu, busy);
        }
        if (busy_m == 0) {
            if (exhausted) break;
            continue;
        }
        /* inner steps: every walker without a pending leaf pops nodes until it finds a leaf */
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (busy && leaf_n == 0) {
                if (sp == 0) {
                    if (unc_whole) unc_emit(unc, id_x, id_y, key, TMK_WHOLE_PAIR);
                    busy = false;
                } else {
                    const BvhNode nd = load_node(nodes, stk[--sp][tid]);
                    if (Q.overlaps(nd)) {
                        if (nd.b > 0) { leaf_a = nd.a; leaf_n = nd.b; }
                        else if (sp + 2 <= TMK_TRAV_STACK) { stk[sp++][tid] = nd.a; stk[sp++][tid] = -nd.b - 1; }
                        else unc_whole = true; /* cannot happen for trees built by the host */
                    }
                }
            }
        }
        /* leaf step: one triangle for every walker that holds a leaf */
        const unsigned pend = __ballot_sync(0xffffffffu, busy && leaf_n > 0);
        const unsigned can_step = __ballot_sync(0xffffffffu, busy && leaf_n == 0);
        if (pend && (__popc(pend) >= 12 || can_step == 0)) {
            if (busy && leaf_n > 0) {
                const int tri = leaf_a;
                leaf_a++; leaf_n--;
                Vec3<float> p0, p1, p2;
                Q.tri_in_s(tris, tri, p0, p1, p2);
                int r;
                if (S.kind == K_SPHERE) r = test_tri_vs_shape(S, p0, p1, p2);
                else r = quick_tri(S, p0, p1, p2);
                if (r == HIT) {
                    sink.mark_hit(src.from_id(id_x, id_y));
                    busy = false; leaf_n = 0;
                } else if (r == NEED) { /* open triangle: to the triangle list (GJK in k_tri_items) */
                    const uint32_t k = atomicAdd(unc.tri_n, 1u);
                    if (k < unc.tri_cap) unc.tri_items[k] = make_uint2(rec, (uint32_t)tri);
                    else unc_emit(unc, id_x, id_y, key, (uint32_t)tri);
                } else if (r == UNC) unc_emit(unc, id_x, id_y, key, (uint32_t)tri);
            }
        }
    }
}

/* The walks of 32 items at a time as two task stacks per warp.  A thread-per-item walk (and the pool above) is as
 * slow as its longest item: a shape lying along the surface visits dozens of nodes and triangles, one dependent
 * load after the other, while the lanes whose walk ended at the root box wait.  Here the warp keeps the queries of
 * its 32 items in shared memory and works through a stack of (item, node) tasks and a stack of (item, triangle)
 * tasks, 32 tasks per round whatever item they belong to: a round is one load latency, the number of rounds is the
 * depth of the tree plus the work divided by 32, and every lane runs the same code (node rounds: box tests;
 * triangle rounds: the per-triangle certificates).  Depth-first in batches (tasks are popped from the end), so the
 * stacks stay short; an item that was hit drops its remaining tasks. */
#define TMK_TASK_NQ 1024
#define TMK_TASK_TQ 320
#define TMK_TASK_F 22 /* floats of a query: Rms 9, tms 3, c 3, extents 3, rad2, shape a b c */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_mesh_tasks(Image<float> im, Src src, Sink sink, UncList unc, uint32_t *next, uint32_t task_cap) {
    __shared__ float s_f[4][TMK_TASK_F][32];
    __shared__ uint32_t s_i[4][6][32]; /* shape kind, mesh record, node pointer (2 words), triangle pointer (2 words) */
    __shared__ uint32_t s_nq[4][TMK_TASK_NQ];
    __shared__ uint32_t s_tq[4][TMK_TASK_TQ];
    __shared__ uint32_t s_dead[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float(*F)[32] = s_f[warp];
    uint32_t(*I)[32] = s_i[warp];
    uint32_t *nq = s_nq[warp], *tq = s_tq[warp];
    const unsigned lt = (1u << lane) - 1u;
    uint32_t n = *unc.mesh_n;
    if (n > unc.mesh_cap) n = unc.mesh_cap;
    /* task_cap: entries of a warp's task stacks it may use (test aid: a small value forces the overflow path).  Not a
     * field of UncList: that struct is passed by value to a noinline function of the tile kernel, and one more word
     * pushed it over the size the ABI passes in registers (tile kernel 93 -> 125 registers, 272 B of stack) */
    const int nq_cap = (int)min(task_cap, (uint32_t)TMK_TASK_NQ), tq_cap = (int)min(task_cap, (uint32_t)TMK_TASK_TQ);
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) { base = atomicAdd(next, 32u); s_dead[warp] = 0u; }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + (uint32_t)lane;
        bool has = false;
        if (i < n) {
            const HeavyRec h = heavy_load(unc.mesh_items + (size_t)i * TMK_MESH_REC);
            const DGeom<float> &A = im.mg[ITEM_A(h.key)];
            const DGeom<float> &B = ITEM_STATIC(h.key) ? im.sg[ITEM_B(h.key)] : im.mg[ITEM_B(h.key)];
            const Xf<float> XB = ITEM_STATIC(h.key) ? xload(B.tf) : h.XB;
            if (A.shape == K_MESH && B.shape == K_MESH) {
                unc_emit(unc, h.id_x, h.id_y, h.key, TMK_WHOLE_PAIR); /* mesh against mesh: exact path only */
            } else {
                const bool a_mesh = A.shape == K_MESH;
                const MeshRef<float> &M = im.meshes[a_mesh ? A.mesh : B.mesh];
                const Shape<float> S = shape_of(a_mesh ? B : A);
                MeshQuery<float> Q;
                Q.init(a_mesh ? h.XA : XB, S, a_mesh ? XB : h.XA, a_mesh ? B.brad : A.brad);
                F[0][lane] = Q.Rms.m00; F[1][lane] = Q.Rms.m01; F[2][lane] = Q.Rms.m02;
                F[3][lane] = Q.Rms.m10; F[4][lane] = Q.Rms.m11; F[5][lane] = Q.Rms.m12;
                F[6][lane] = Q.Rms.m20; F[7][lane] = Q.Rms.m21; F[8][lane] = Q.Rms.m22;
                F[9][lane] = Q.tms.x; F[10][lane] = Q.tms.y; F[11][lane] = Q.tms.z;
                F[12][lane] = Q.c.x; F[13][lane] = Q.c.y; F[14][lane] = Q.c.z;
                F[15][lane] = Q.ex; F[16][lane] = Q.ey; F[17][lane] = Q.ez; F[18][lane] = Q.rad2;
                F[19][lane] = S.a; F[20][lane] = S.b; F[21][lane] = S.c;
                I[0][lane] = (uint32_t)S.kind; I[1][lane] = i;
                const unsigned long long pn = (unsigned long long)(uintptr_t)M.nodes, pt = (unsigned long long)(uintptr_t)M.tris;
                I[2][lane] = (uint32_t)pn; I[3][lane] = (uint32_t)(pn >> 32);
                I[4][lane] = (uint32_t)pt; I[5][lane] = (uint32_t)(pt >> 32);
                has = true;
            }
        }
        const unsigned hm = __ballot_sync(0xffffffffu, has);
        if (has) nq[__popc(hm & lt)] = (uint32_t)lane << 24; /* the root of every item */
        int nq_n = __popc(hm), tq_n = 0;
        bool overflow = false;
        __syncwarp();
        while (nq_n > 0 || tq_n > 0) {
            const uint32_t dead = s_dead[warp];
            if (tq_n >= 32 || nq_n == 0) { /* a round of triangles */
                const int take = min(32, tq_n);
                tq_n -= take;
                if (lane < take) {
                    const uint32_t t = tq[tq_n + lane];
                    const int sl = (int)(t >> 24), tri = (int)(t & 0xffffffu);
                    if (!((dead >> sl) & 1u)) {
                        MeshQuery<float> Q;
                        Q.Rms.m00 = F[0][sl]; Q.Rms.m01 = F[1][sl]; Q.Rms.m02 = F[2][sl];
                        Q.Rms.m10 = F[3][sl]; Q.Rms.m11 = F[4][sl]; Q.Rms.m12 = F[5][sl];
                        Q.Rms.m20 = F[6][sl]; Q.Rms.m21 = F[7][sl]; Q.Rms.m22 = F[8][sl];
                        Q.tms = mk<float>(F[9][sl], F[10][sl], F[11][sl]);
                        Shape<float> S;
                        S.kind = (int)I[0][sl]; S.a = F[19][sl]; S.b = F[20][sl]; S.c = F[21][sl];
                        const float *tris = (const float *)(uintptr_t)((unsigned long long)I[4][sl] | ((unsigned long long)I[5][sl] << 32));
                        Vec3<float> p0, p1, p2;
                        Q.tri_in_s(tris, tri, p0, p1, p2);
                        int r;
                        if (S.kind == K_SPHERE) r = test_tri_vs_shape(S, p0, p1, p2);
                        else r = quick_tri(S, p0, p1, p2);
                        if (r != SEP) {
                            const uint32_t rec = I[1][sl];
                            if (r == NEED) { /* open triangle: to the triangle list (GJK in k_tri_items) */
                                const uint32_t k = atomicAdd(unc.tri_n, 1u);
                                if (k < unc.tri_cap) unc.tri_items[k] = make_uint2(rec, (uint32_t)tri);
                                else r = UNC;
                            }
                            if (r == HIT || r == UNC) {
                                const float4 r0 = unc.mesh_items[(size_t)rec * TMK_MESH_REC];
                                if (r == HIT) {
                                    sink.mark_hit(src.from_id(__float_as_uint(r0.x), __float_as_uint(r0.y)));
                                    atomicOr(&s_dead[warp], 1u << sl);
                                } else unc_emit(unc, __float_as_uint(r0.x), __float_as_uint(r0.y), __float_as_uint(r0.z), (uint32_t)tri);
                            }
                        }
                    }
                }
            } else { /* a round of nodes */
                const int take = min(32, nq_n);
                nq_n -= take;
                int ca = 0, cb = 0, sl = 0; /* what this lane pushes: two children (cb < 0) or cb triangles from ca on */
                if (lane < take) {
                    const uint32_t t = nq[nq_n + lane];
                    sl = (int)(t >> 24);
                    if (!((dead >> sl) & 1u)) {
                        const BvhNode *nodes = (const BvhNode *)(uintptr_t)((unsigned long long)I[2][sl] | ((unsigned long long)I[3][sl] << 32));
                        const BvhNode nd = load_node(nodes, (int)(t & 0xffffffu));
                        MeshQuery<float> Q;
                        Q.c = mk<float>(F[12][sl], F[13][sl], F[14][sl]);
                        Q.ex = F[15][sl]; Q.ey = F[16][sl]; Q.ez = F[17][sl]; Q.rad2 = F[18][sl];
                        if (Q.overlaps(nd)) { ca = nd.a; cb = nd.b; }
                    }
                }
                __syncwarp(); /* the popped entries were read: the slots may be overwritten */
                const unsigned inner = __ballot_sync(0xffffffffu, cb < 0);
                const int ntri = cb > 0 ? cb : 0;
                /* slots of this lane's triangles: the host's trees have leaves of one or two triangles - two ballots;
                 * larger leaves take the shuffle scan */
                const unsigned m1 = __ballot_sync(0xffffffffu, ntri >= 1), m2 = __ballot_sync(0xffffffffu, ntri >= 2);
                int excl = __popc(m1 & lt) + __popc(m2 & lt), tri_total = __popc(m1) + __popc(m2);
                if (__any_sync(0xffffffffu, ntri > 2)) {
                    int incl = ntri;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    tri_total = __shfl_sync(0xffffffffu, incl, 31);
                    excl = incl - ntri;
                }
                if (nq_n + 2 * __popc(inner) > nq_cap || tq_n + tri_total > tq_cap) { overflow = true; break; }
                if (cb < 0) {
                    const int o = nq_n + 2 * __popc(inner & lt);
                    nq[o] = ((uint32_t)sl << 24) | (uint32_t)ca;
                    nq[o + 1] = ((uint32_t)sl << 24) | (uint32_t)(-cb - 1);
                }
                const uint32_t t0 = ((uint32_t)sl << 24) | (uint32_t)ca;
                if (ntri >= 1) tq[tq_n + excl] = t0;
                if (ntri >= 2) tq[tq_n + excl + 1] = t0 + 1u;
                for (int j = 2; j < ntri; j++) tq[tq_n + excl + j] = t0 + (uint32_t)j;
                nq_n += 2 * __popc(inner);
                tq_n += tri_total;
            }
            __syncwarp();
        }
        if (overflow) { /* cannot happen for the trees the host builds: every item not yet hit goes to the exact path */
            if (has && !((s_dead[warp] >> lane) & 1u)) {
                const float4 r0 = unc.mesh_items[(size_t)i * TMK_MESH_REC];
                unc_emit(unc, __float_as_uint(r0.x), __float_as_uint(r0.y), __float_as_uint(r0.z), TMK_WHOLE_PAIR);
            }
        }
        __syncwarp();
    }
}

/* both lists in one launch, triangles decided inline: for the swept-edge levels, whose lists are short
 * and whose launch count matters more than warp convergence */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_heavy_mixed(Image<float> im, Src src, Sink sink, UncList unc) {
    uint32_t nc = *unc.conv_n, nm = *unc.mesh_n;
    if (nc > unc.conv_cap) nc = unc.conv_cap;
    if (nm > unc.mesh_cap) nm = unc.mesh_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nc + nm; i += gridDim.x * blockDim.x) {
        const bool conv = i < nc;
        const HeavyRec h = heavy_load((conv ? unc.conv_items : unc.mesh_items) + (size_t)(conv ? i : i - nc) * TMK_MESH_REC);
        const DGeom<float> &A = im.mg[ITEM_A(h.key)];
        const DGeom<float> &B = ITEM_STATIC(h.key) ? im.sg[ITEM_B(h.key)] : im.mg[ITEM_B(h.key)];
        const Xf<float> XB = ITEM_STATIC(h.key) ? xload(B.tf) : h.XB;
        TriEmit on_unc;
        on_unc.u = unc; on_unc.id_x = h.id_x; on_unc.id_y = h.id_y; on_unc.pair = h.key;
        const int r = test_geoms(A, h.XA, B, XB, im.meshes, on_unc);
        if (r == HIT) sink.mark_hit(src.from_id(h.id_x, h.id_y));
        else if (r == UNC) unc_emit(unc, h.id_x, h.id_y, h.key, TMK_WHOLE_PAIR);
    }
}

/* one open triangle against the pair's convex shape: GJK */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_tri_items(Image<float> im, Src src, Sink sink, UncList unc) {
    uint32_t n = *unc.tri_n;
    if (n > unc.tri_cap) n = unc.tri_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 it = unc.tri_items[i];
        const HeavyRec h = heavy_load(unc.mesh_items + (size_t)it.x * TMK_MESH_REC);
        const DGeom<float> &A = im.mg[ITEM_A(h.key)];
        const DGeom<float> &B = ITEM_STATIC(h.key) ? im.sg[ITEM_B(h.key)] : im.mg[ITEM_B(h.key)];
        const Xf<float> XB = ITEM_STATIC(h.key) ? xload(B.tf) : h.XB;
        int r;
        if (A.shape == K_MESH) r = test_mesh_tri_shape(im.meshes[A.mesh], h.XA, shape_of(B), XB, (int)it.y);
        else r = test_mesh_tri_shape(im.meshes[B.mesh], XB, shape_of(A), h.XA, (int)it.y);
        if (r == HIT) sink.mark_hit(src.from_id(h.id_x, h.id_y));
        else if (r == UNC) unc_emit(unc, h.id_x, h.id_y, h.key, it.y);
    }
}

/* exact re-check of the (state, pair) items the FP32 certificates could not decide */
template <class Src, class Sink>
__global__ void __launch_bounds__(128) k_recheck_items(Image<double> im, Src src, Sink sink, UncList unc) {
    uint32_t n = *unc.n;
    if (n > unc.cap) n = unc.cap;
    /* The list is short (about a thousand items per 2^20 states) and every item takes its own path through the
     * double-precision code (GJK of different shapes, iteration counts, triangles): 32 of them in one warp run
     * one after the other.  So the items are spread as thinly as the grid allows - one per warp when there are
     * fewer items than warps - and only packed when the list is long. */
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t warp_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t ipw = min(32u, (n + n_warps - 1) / n_warps); /* items per warp and round */
    for (uint32_t base = warp_g * ipw; base < n; base += n_warps * ipw) {
        const uint32_t i = base + lane;
        if (lane >= ipw || i >= n) continue;
        const uint4 it = unc.items[i];
        const uint32_t pair = it.z;
        const typename Src::State st = src.from_id(it.x, it.y);
        if (!st.live) continue;
        double q[TMK_MAXSUB];
        for (int k = 0; k < im.n_sub; k++) q[k] = src.template load1<double>(st, k);
        const int a = ITEM_A(pair), b = ITEM_B(pair);
        const DGeom<double> &A = im.mg[a];
        const DGeom<double> &B = ITEM_STATIC(pair) ? im.sg[b] : im.mg[b];
        Xf<double> XA, XB;
        fk_pair(im, q, a, ITEM_STATIC(pair) ? -1 : b, XA, XB); /* only the joints these two geometries hang on */
        if (ITEM_STATIC(pair)) XB = xload(B.tf);
        int r;
        if (it.w == TMK_WHOLE_PAIR) r = test_geoms(A, XA, B, XB, im.meshes);
        else if (A.shape == K_MESH) r = test_mesh_tri_shape(im.meshes[A.mesh], XA, shape_of(B), XB, (int)it.w);
        else r = test_mesh_tri_shape(im.meshes[B.mesh], XB, shape_of(A), XA, (int)it.w);
        if (r == HIT) sink.mark_hit(st);
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
/* An edge with more segments than the launched bisection levels cover (positions 0 .. max_nd - 1) cannot be
 * decided level by level: it raises the overflow flag and k_edges_overflow_all walks every edge's states in
 * OMPL order instead (slow, exact, and only for a seg_len far below the reference's 1 % resolution). */
__global__ void k_edges_prep(const void *p0, const void *p1, int layout, size_t ld, int n_sub, uint32_t n_edge,
                             double seg_len, int32_t *nd, int32_t *first, int static_hit, int max_nd,
                             uint32_t *overflow) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edge) return;
    double a[TMK_MAXSUB], b[TMK_MAXSUB];
    for (int k = 0; k < n_sub; k++) {
        a[k] = load_q<double>(p0, layout, ld, e, k);
        b[k] = load_q<double>(p1, layout, ld, e, k);
    }
    const int n = segment_count(a, b, n_sub, seg_len);
    nd[e] = n;
    first[e] = static_hit ? 0 : TMK_FIRST_NONE;
    if (n > max_nd) *overflow = 1u;
}

/* segment counts, first[] reset and the exclusive prefix sum of the counts for SrcEdgeFlat: one CTA, n_edge <= 4096 */
__global__ void __launch_bounds__(1024) k_edges_prep_flat(const void *p0, const void *p1, int layout, size_t ld, int n_sub, uint32_t n_edge,
                                                         double seg_len, int32_t *nd, int32_t *first, int static_hit, uint32_t *off) {
    __shared__ uint32_t warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t mine[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t e = 4u * tid + k;
        mine[k] = 0;
        if (e < n_edge) {
            double a[TMK_MAXSUB], b[TMK_MAXSUB];
            for (int j = 0; j < n_sub; j++) {
                a[j] = load_q<double>(p0, layout, ld, e, j);
                b[j] = load_q<double>(p1, layout, ld, e, j);
            }
            const int n = segment_count(a, b, n_sub, seg_len);
            nd[e] = n;
            first[e] = static_hit ? 0 : TMK_FIRST_NONE;
            mine[k] = (uint32_t)n; /* positions 0 .. nd-1: nd states */
        }
        sum += mine[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        warp_sum[lane] = wi - w; /* exclusive */
    }
    __syncthreads();
    uint32_t run = warp_sum[warp] + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t e = 4u * tid + k;
        if (e < n_edge) off[e] = run;
        run += mine[k];
        if (e + 1 == n_edge) off[n_edge] = run;
    }
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

#define TMK_STAGE_BUDGET (48 * 1024) /* shared memory the staged part of the image may take */

/* shared-memory plan for one (kernel, tile) choice; returns resident warps per SM (0 = does not fit) */
template <class Kernel>
static int tile_plan(TileCfg &cfg, const Image<float> &im, int n_jslab, Kernel kernel, int tile, int sms, int optin,
                     int sm_total) {
    cfg.tile = tile;
    cfg.bytes_joint = align16(im.n_joint * (int)sizeof(DJoint<float>));
    cfg.bytes_mg = align16(im.n_mg * (int)sizeof(DGeom<float>));
    cfg.bytes_sg = align16(im.n_sg * (int)sizeof(DGeom<float>));
    cfg.bytes_pairs = align16(im.n_tpair * 4);
    cfg.bytes_sq = align16(im.n_sg * (int)sizeof(SQuick));
    cfg.bytes_jg = align16((im.n_joint + 1) * 4);
    cfg.bytes_pa = align16((3 * im.n_mg + 1) * 4);
    cfg.bytes_ps4 = im.n_tpair * 16;
    /* the static grid's walk arrays: cell starts, item ids, item spheres in cell order */
    cfg.bytes_gcell = im.grid.on ? align16((im.grid.nx * im.grid.ny * im.grid.nz + 1) * 4) : 0;
    cfg.bytes_gitems = im.grid.on ? align16((im.grid.n_items + 1) * 2) : 0;
    cfg.bytes_gc4 = im.grid.on ? (im.grid.n_items + 1) * 16 : 0;
    /* staging policy: everything if it fits the budget; else leave the static geometry table (only read by
     * the narrowphase phases) in global memory; else the quick table too; else nothing */
    {
        int small = cfg.bytes_joint + cfg.bytes_mg + cfg.bytes_jg + cfg.bytes_pa;
        /* the grid walk is the inner loop of a large scene: its arrays go first, the quick tables (read once per
         * surviving pair) last */
        if (small + cfg.bytes_gcell + cfg.bytes_gitems + cfg.bytes_gc4 <= TMK_STAGE_BUDGET) small += cfg.bytes_gcell + cfg.bytes_gitems + cfg.bytes_gc4;
        else cfg.bytes_gcell = cfg.bytes_gitems = cfg.bytes_gc4 = 0;
        const int lists = cfg.bytes_pairs + cfg.bytes_ps4;
        cfg.stage = 1;
        if (small + lists + cfg.bytes_sq + cfg.bytes_sg <= TMK_STAGE_BUDGET) { /* all */
        } else if (small + lists + cfg.bytes_sq <= TMK_STAGE_BUDGET) cfg.bytes_sg = 0;
        else if (small + lists <= TMK_STAGE_BUDGET) cfg.bytes_sg = cfg.bytes_sq = 0;
        else if (small <= TMK_STAGE_BUDGET) cfg.bytes_sg = cfg.bytes_sq = cfg.bytes_pairs = cfg.bytes_ps4 = 0;
        else cfg.stage = 0;
        if (!cfg.stage) cfg.bytes_gcell = cfg.bytes_gitems = cfg.bytes_gc4 = 0;
    }
    int off = 0;
    cfg.off_joint = off; off += cfg.stage ? cfg.bytes_joint : 0;
    cfg.off_mg = off; off += cfg.stage ? cfg.bytes_mg : 0;
    cfg.off_sg = off; off += cfg.stage ? cfg.bytes_sg : 0;
    cfg.off_pairs = off; off += cfg.stage ? cfg.bytes_pairs : 0;
    cfg.off_sq = off; off += cfg.stage ? cfg.bytes_sq : 0;
    cfg.off_jg = off; off += cfg.stage ? cfg.bytes_jg : 0;
    cfg.off_pa = off; off += cfg.stage ? cfg.bytes_pa : 0;
    cfg.off_ps4 = off; off += cfg.stage ? cfg.bytes_ps4 : 0;
    cfg.off_gcell = off; off += cfg.bytes_gcell;
    cfg.off_gitems = off; off += cfg.bytes_gitems;
    cfg.off_gc4 = off; off += cfg.bytes_gc4;
    off = (off + 127) & ~127;
    cfg.off_slab = off; off += im.n_mg * 7 * tile * 4;
    cfg.off_qs = off; off += im.n_sub * tile * 4;
    cfg.off_jslab = off; off += n_jslab * 7 * tile * 4;
    cfg.off_flags = off; off += tile;
    cfg.off_unc = off; off += 2 * tile * 4;
    off = (off + 15) & ~15;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 0;
    int max_dyn = optin - (int)fa.sharedSizeBytes;
    /* queues share what is left of the per-CTA budget at the best occupancy the fixed part allows */
    const int per_cta_over = 1024 + (int)fa.sharedSizeBytes; /* driver reservation + static */
    int best_ctas = 0, reg_ctas = 2048 / tile; /* at most 64 warps per SM; fewer if registers say so */
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&reg_ctas, kernel, tile, 0) != cudaSuccess || reg_ctas < 1) return 0;
    for (int ctas = reg_ctas; ctas >= 1; ctas--) {
        int budget = sm_total / ctas - per_cta_over;
        if (budget > max_dyn) budget = max_dyn;
        int left = budget - off;
        if (left >= 12 * 1024 || (ctas == 1 && left >= 4 * 1024)) { best_ctas = ctas; break; }
    }
    if (!best_ctas) return 0;
    int budget = sm_total / best_ctas - per_cta_over;
    if (budget > max_dyn) budget = max_dyn;
    int left = (budget - off) & ~15;
    /* q1 / q3 are split into one private region per warp (capacities below are PER WARP);
     * 9/16 broadphase survivors, 2/16 convex narrowphase, 5/16 mesh narrowphase */
    const int nw = tile / 32;
    cfg.q2_cap = left / 8 / 4;
    cfg.q3_cap = (left * 5 / 16) / 4 / nw;
    cfg.q1_cap = (left - 4 * cfg.q2_cap - 4 * cfg.q3_cap * nw) / 4 / nw;
    if (cfg.q1_cap > (1 << 16)) cfg.q1_cap = 1 << 16;
    if (const char *e = getenv("TMK_QCAP")) { /* test aid: tiny queues exercise the overflow -> exact re-check path */
        int c = atoi(e);
        if (c > 0) { cfg.q1_cap = std::min(cfg.q1_cap, c); cfg.q2_cap = std::min(cfg.q2_cap, c); cfg.q3_cap = std::min(cfg.q3_cap, c); }
    }
    cfg.off_q1 = off; off += cfg.q1_cap * nw * 4;
    cfg.off_q2 = off; off += cfg.q2_cap * 4;
    cfg.off_q3 = off; off += cfg.q3_cap * nw * 4;
    cfg.smem = off;
    /* the attribute is per function, not per collision context: always raise it to the maximum */
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) != cudaSuccess) return 0;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, tile, cfg.smem);
    if (per_sm < 1) return 0;
    cfg.grid = sms * per_sm;
    return per_sm * tile / 32;
}

template <class K256, class K128>
static void tile_configure(TileCfg &cfg, const Image<float> &im, int n_jslab, K256 k256, K128 k128) {
    int dev = 0, sms = 148, optin = 227 * 1024, sm_total = 228 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sm_total, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    TileCfg c256, c128;
    int w256 = tile_plan(c256, im, n_jslab, k256, 256, sms, optin, sm_total);
    int w128 = tile_plan(c128, im, n_jslab, k128, 128, sms, optin, sm_total);
    if (!w256 && !w128) { cfg.smem = -1; cfg.grid = 0; cfg.tile = 0; return; }
    const char *force = getenv("TMK_TILE"); /* tuning aid */
    if (force && atoi(force) == 128 && w128) { cfg = c128; return; }
    if (force && atoi(force) == 256 && w256) { cfg = c256; return; }
    cfg = (w256 >= w128) ? c256 : c128;
}

} /* namespace tmk */
