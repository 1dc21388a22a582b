/* wave.cuh - the state-validity pipeline for configuration batches, wavefront style (BASELINE.json configs[1], [3],
 * [4]: every batch of 4096 states or more; the swept-edge levels and small batches stay on the tile kernel).
 *
 * The tile kernel (pipeline.cuh) keeps every pose of a state in shared memory so that nothing touches HBM: two CTAs
 * per SM for the 10 geometries of one arm (issue slots 53 % busy), one CTA for the 20 of two arms (24 %).  The
 * arithmetic per state (~11 k instructions) dwarfs the bytes of its poses (280 - 560 B), so here the phases are
 * separate kernels that hand their data through HBM / L2 and each runs at full occupancy:
 *
 *   k_wave_fk      one thread = one state: the joint chain in registers, every moving geometry's pose written to
 *                  planes [geom * 7 + k][state] (coalesced), hit flag cleared, moving-moving pairs voted
 *   k_wave_broad   one thread = K = 2 states of one moving geometry (a0 + blockIdx.y), so the partner lists are
 *                  CTA-uniform (broadcast loads) and shared by the two tests of a pair: bounding spheres against a
 *                  precomputed squared reach, world-aligned boxes as centre + half extents, oriented boxes exactly;
 *                  the small obstacles of large scenes come from the reach list of the one grid cell the geometry's
 *                  centre lies in (a contiguous run of half-precision spheres, rounded conservatively, read
 *                  through L2); survivors are staged per warp in shared memory (the fill is a warp-uniform
 *                  register) and flushed with one global atomic per CTA.  Run in stages over index ranges of the
 *                  geometries, each followed by its k_wave_quick: a state a stage condemns skips the later ones
 *   k_wave_quick   one thread = one surviving (state, pair): quick certificates (dev_math.cuh quick_pair); HIT sets
 *                  the state's flag, undecided pairs go to a (state, pair) list, FP32-ambiguous ones to the exact
 *                  re-check list
 *   k_wave_export  one thread = one undecided pair of a state that is still not known to collide: the record the
 *                  deferred narrowphase kernels of pipeline.cuh consume (k_conv_items, k_mesh_tasks, k_tri_items)
 *   k_wave_bits    one thread = one state: the validity words
 *
 * Verdicts are those of the tile kernel bit for bit: the same tests in the same arithmetic decide every pair.
 * No reference counterpart exists (the reference checks one state at a time on the CPU).
 */
#pragma once
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include "pipeline.cuh"

namespace tmk {

struct WaveBuf {
    float *pose;        /* [(geom * 7 + k) * cap + state]: x y z w of the rotation, then the translation */
    uint8_t *hit;       /* [cap] 1 = the state is known to collide */
    uint2 *queue;       /* broadphase survivors: (state in the chunk, pair key = ITEM_KEY layout) */
    uint2 *pend;        /* pairs the quick certificates left open */
    uint32_t *qcnt;     /* queue fill of the stage in flight */
    uint32_t *pcnt;     /* pend fill (all stages of the chunk) */
    uint32_t cap;       /* states per chunk */
    uint32_t queue_cap, pend_cap;
};
#define TMK_WAVE_STAGE 224 /* survivors staged per warp before a flush */

__device__ __forceinline__ Xf<float> wave_pose(const WaveBuf &w, int g, uint32_t s) {
    const float *o = w.pose + (size_t)g * 7 * w.cap + s;
    Xf<float> X;
    X.r.x = o[0]; X.r.y = o[w.cap]; X.r.z = o[2 * (size_t)w.cap]; X.r.w = o[3 * (size_t)w.cap];
    X.v.x = o[4 * (size_t)w.cap]; X.v.y = o[5 * (size_t)w.cap]; X.v.z = o[6 * (size_t)w.cap];
    return X;
}

/* warp-level staging of broadphase survivors: appended in shared memory, flushed to the queue with one atomic */
struct WaveStage {
    uint2 *stage;      /* this warp's slots */
    uint32_t cap = TMK_WAVE_STAGE; /* ... how many */
    uint32_t *cnt;     /* this warp's fill */
    __device__ __forceinline__ void put(const WaveBuf &w, const UncList &unc, uint32_t idx, uint32_t s, int a, int b, int b_static) const {
        put_key(w, unc, idx, s, ITEM_KEY(item_pack(0, a, b, b_static)));
    }
    __device__ __forceinline__ void put_key(const WaveBuf &w, const UncList &unc, uint32_t idx, uint32_t s, uint32_t key) const {
        const uint2 item = make_uint2(s, key);
        if (idx < cap) stage[idx] = item;
        else { /* the warp's staging area is full: straight to the queue */
            const uint32_t k = atomicAdd(w.qcnt, 1u);
            if (k < w.queue_cap) w.queue[k] = item;
            else *unc.overflow = 1u;
        }
    }
    /* warp-uniform calls: every lane votes, passing lanes get consecutive slots.  The warp's fill is a warp-uniform
     * register (every lane adds the same popc): no atomic, no shuffle; publish() stores it for the flush */
    uint32_t fill;
    /* key: the pair's queue key, ready to store (Image::t_keys) */
    __device__ __forceinline__ void vote_key(const WaveBuf &w, const UncList &unc, bool pass, uint32_t s, uint32_t key, int lane) {
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m == 0u) return; /* warp-uniform: most votes of a group that had a survivor somewhere are empty */
        if (pass) put_key(w, unc, fill + __popc(m & ((1u << lane) - 1u)), s, key);
        fill += __popc(m);
    }
    __device__ __forceinline__ void publish(int lane) const {
        if (lane == 0) *cnt = fill;
        __syncwarp();
    }
    /* CTA-wide flush (every thread of the CTA calls it; `stage`/`cnt` of warp 0 must be passed as stage0/cnt0): one
     * global atomic per CTA - the queue counter is a single address every CTA of the launch appends to */
    __device__ __forceinline__ static void flush_cta(const WaveBuf &w, const UncList &unc, const uint2 *stage0, const uint32_t *cnt0,
                                                     uint32_t *s_base, int tid, uint32_t stage_cap = TMK_WAVE_STAGE) {
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        /* fills of the eight warps: lane k holds warp k's, a shuffle scan gives the total and this warp's offset */
        uint32_t f = lane < 8 ? min(cnt0[lane], stage_cap) : 0u, incl = f;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 7);
        const uint32_t mine = __shfl_sync(0xffffffffu, f, warp), before = __shfl_sync(0xffffffffu, incl, warp) - mine;
        if (total == 0) return; /* CTA-uniform */
        if (tid == 0) *s_base = atomicAdd(w.qcnt, total);
        __syncthreads();
        const uint32_t base = *s_base + before; /* every warp copies its own region */
        const uint2 *src = stage0 + (size_t)warp * stage_cap;
        for (uint32_t i = lane; i < mine; i += 32) {
            if (base + i < w.queue_cap) w.queue[base + i] = src[i];
            else *unc.overflow = 1u;
        }
    }
};

/* FK of every state + the broadphase of the MOVING partners: when a geometry's pose exists the centres of the
 * moving geometries before it are still in the thread's hands (a pair of two moving geometries is owned by the
 * later one), so those pairs are voted on here instead of re-reading partner centres from HBM per pair */
template <class Src>
__global__ void __launch_bounds__(256) k_wave_fk(Image<float> im, Src src, WaveBuf w, UncList unc) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t cnt[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WaveStage ws;
    ws.stage = (uint2 *)smem + (size_t)warp * TMK_WAVE_STAGE;
    ws.cnt = &cnt[warp];
    ws.fill = 0;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const typename Src::State st = src.prepare(s);
    const bool live = st.live && !im.static_hit;
    if (st.live) w.hit[s] = im.static_hit ? 1 : 0;
    const float pad = Num<float>::eps() * 4.0f;
    /* the chain runs in registers; only joints with a child that does not follow them directly (pad != 0) are kept in
     * the per-thread array (dynamically indexed: local memory) */
    Xf<float> jp[TMK_MAXJ];
    /* bounding centres of the moving geometries seen so far, indexed by the partner: local memory (the same table in
     * shared memory costs occupancy and is slower: configs 0.723 -> 0.739 ms, 2^24 clutter 20.9 -> 24.6 ms) */
    float gx[TMK_MAXMG], gy[TMK_MAXMG], gz[TMK_MAXMG];
    Xf<float> cur;
    cur.r.x = cur.r.y = cur.r.z = 0.f; cur.r.w = 1.f; cur.v = mk<float>(0.f, 0.f, 0.f);
    for (int j = 0; j < im.n_joint; j++) {
        const DJoint<float> &J = im.joints[j];
        Xf<float> par = cur;
        if (J.parent >= 0 && J.parent != j - 1) par = jp[J.parent];
        const float qv = st.live ? src.template load1<float>(st, J.qslot) : 0.f;
        cur = fk_joint(xload(J.E), mk<float>(J.ax[0], J.ax[1], J.ax[2]), J.off, J.type, qv,
                       J.parent >= 0 ? &par : (const Xf<float> *)nullptr);
        if (J.pad) jp[j] = cur;
        for (int a = im.jg_begin[j]; a < im.jg_begin[j + 1]; a++) {
            const DGeom<float> &A = im.mg[a];
            const Xf<float> XA = xmul(cur, xload(A.tf));
            if (st.live) {
                float *o = w.pose + (size_t)a * 7 * w.cap + s;
                o[0] = XA.r.x; o[w.cap] = XA.r.y; o[2 * (size_t)w.cap] = XA.r.z; o[3 * (size_t)w.cap] = XA.r.w;
                o[4 * (size_t)w.cap] = XA.v.x; o[5 * (size_t)w.cap] = XA.v.y; o[6 * (size_t)w.cap] = XA.v.z;
            }
            const Vec3<float> ca = bound_centre(A, XA);
            gx[a] = ca.x; gy[a] = ca.y; gz[a] = ca.z;
            /* moving partners: the partner's bounding radius rides in ps4[p].w; four pairs per vote (most groups
             * have no survivor in any lane), their loads issued together */
            const uint32_t *PAIRS = im.t_pairs;
            const float4 *PS4 = (const float4 *)im.ps4;
            int p = im.pa_begin[3 * a + 2];
            const int pe = im.pa_begin[3 * a + 3];
            for (; p + 4 <= pe; p += 4) {
                int b[4];
                bool pass[4];
#pragma unroll
                for (int u = 0; u < 4; u++) b[u] = (int)((PAIRS[p + u] >> 8) & 0xffff);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float dx = gx[b[u]] - ca.x, dy = gy[b[u]] - ca.y, dz = gz[b[u]] - ca.z;
                    pass[u] = live && dx * dx + dy * dy + dz * dz <= PS4[p + u].w; /* .w = (ra + pad + rb)^2 */
                }
                if (__any_sync(0xffffffffu, pass[0] | pass[1] | pass[2] | pass[3])) {
#pragma unroll
                    for (int u = 0; u < 4; u++) ws.vote_key(w, unc, pass[u], s, im.t_keys[p + u], lane);
                }
            }
            for (; p < pe; p++) {
                const int b = (int)((PAIRS[p] >> 8) & 0xffff);
                const float dx = gx[b] - ca.x, dy = gy[b] - ca.y, dz = gz[b] - ca.z;
                ws.vote_key(w, unc, live && dx * dx + dy * dy + dz * dz <= PS4[p].w, s, im.t_keys[p], lane);
            }
        }
    }
    ws.publish(lane);
    WaveStage::flush_cta(w, unc, (const uint2 *)smem, cnt, &s_base, tid);
}

/* broadphase of moving geometry a = a0 + blockIdx.y against its partner lists and the static grid.
 * a0: first geometry of the stage (the index stays in the uniform datapath - read from a table it would not, and
 * every partner-list load of the kernel would hang off a local-memory load: +15 %).
 * K states per thread (s, s + 256, ...): the partner data, the loop and its address arithmetic are shared by the K
 * tests of a pair. */
template <int K>
__global__ void __launch_bounds__(256) k_wave_broad(Image<float> im, WaveBuf w, uint32_t n, UncList unc, size_t id_base, int a0) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t cnt[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const StaticGrid &G = im.grid;
    constexpr uint32_t STAGE = K > 2 ? 2 * TMK_WAVE_STAGE : TMK_WAVE_STAGE; /* survivors staged per warp */
    uint2 *stage = (uint2 *)smem + (size_t)warp * STAGE;

    const int a = a0 + (int)blockIdx.y;
    const DGeom<float> &A = im.mg[a];
    const size_t cap = w.cap;
    uint32_t s[K];
    bool live[K];
    Vec3<float> ca[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        s[k] = blockIdx.x * (256u * K) + 256u * k + tid;
        live[k] = s[k] < n && w.hit[s[k]] == 0;
        const uint32_t sc = s[k] < n ? s[k] : 0;
        const float *oa = w.pose + (size_t)a * 7 * cap + sc;
        ca[k] = mk<float>(oa[4 * cap], oa[5 * cap], oa[6 * cap]);
        if (A.shape == K_MESH) ca[k] = bound_centre(A, wave_pose(w, a, sc));
        if (!live[k]) ca[k] = mk<float>(1.0e18f, 1.0e18f, 1.0e18f); /* no test passes: squared distances ~ 1e36, no overflow */
    }
    const float ra = A.brad + Num<float>::eps() * 4.0f, ra2 = ra * ra;
    const uint32_t *PAIRS = im.t_pairs;
    const float4 *PS4 = (const float4 *)im.ps4;
    const SQuick *SQ = (const SQuick *)im.sq;
    const int *PA = im.pa_begin;

    WaveStage ws;
    ws.stage = stage;
    ws.cap = STAGE;
    ws.cnt = &cnt[warp];
    ws.fill = 0;
    const uint32_t *KEYS = im.t_keys;
    auto enqueue = [&](int p, const bool (&pass)[K]) { /* warp-uniform call */
        const uint32_t key = KEYS[p];
#pragma unroll
        for (int k = 0; k < K; k++) ws.vote_key(w, unc, pass[k], s[k], key, lane);
    };
    int p = PA[3 * a];
    const int e0 = PA[3 * a + 1], e1 = PA[3 * a + 2], e2 = PA[3 * a + 3];
    /* static partners by bounding sphere: four tests, one vote (most groups have no survivor in any lane) */
    for (; p + 4 <= e0; p += 4) {
        bool pass[4][K];
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float4 c4 = PS4[p + u];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const float dx = c4.x - ca[k].x, dy = c4.y - ca[k].y, dz = c4.z - ca[k].z;
                pass[u][k] = dx * dx + dy * dy + dz * dz <= c4.w; /* .w = (ra + pad + rb)^2; a dead lane's centre is far away */
                any |= pass[u][k];
            }
        }
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int u = 0; u < 4; u++) enqueue(p + u, pass[u]);
        }
    }
    for (; p < e0; p++) {
        const float4 c4 = PS4[p];
        bool pass[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float dx = c4.x - ca[k].x, dy = c4.y - ca[k].y, dz = c4.z - ca[k].z;
            pass[k] = dx * dx + dy * dy + dz * dz <= c4.w;
        }
        enqueue(p, pass);
    }
    /* static boxes / mesh boxes: bounding sphere of A against the box itself (exact point-box distance) */
    auto box_pass = [&](int pp, bool (&pass)[K]) {
        const int b = (PAIRS[pp] >> 8) & 0xffff;
        const float4 c4 = *reinterpret_cast<const float4 *>(&SQ[b].cx);
        const float4 x4 = *reinterpret_cast<const float4 *>(&SQ[b].a0x);
        const float4 y4 = *reinterpret_cast<const float4 *>(&SQ[b].a1x);
        const float4 z4 = *reinterpret_cast<const float4 *>(&SQ[b].a2x);
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float dx = c4.x - ca[k].x, dy = c4.y - ca[k].y, dz = c4.z - ca[k].z;
            const float ex = fmaxf(fabsf(dx * x4.x + dy * x4.y + dz * x4.z) - x4.w, 0.f);
            const float ey = fmaxf(fabsf(dx * y4.x + dy * y4.y + dz * y4.z) - y4.w, 0.f);
            const float ez = fmaxf(fabsf(dx * z4.x + dy * z4.y + dz * z4.z) - z4.w, 0.f);
            pass[k] = ex * ex + ey * ey + ez * ez <= ra2;
            any |= pass[k];
        }
        return any;
    };
    /* the world-aligned boxes of the segment first (tables, walls, most mesh boxes): centre and half extents, no axes */
    const float4 *AB4 = (const float4 *)im.ab4;
    auto aabb_pass = [&](int pp, bool (&pass)[K]) {
        const float4 c4 = AB4[2 * pp], h4 = AB4[2 * pp + 1];
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float ex = fmaxf(fabsf(c4.x - ca[k].x) - h4.x, 0.f), ey = fmaxf(fabsf(c4.y - ca[k].y) - h4.y, 0.f),
                        ez = fmaxf(fabsf(c4.z - ca[k].z) - h4.z, 0.f);
            pass[k] = ex * ex + ey * ey + ez * ez <= ra2;
            any |= pass[k];
        }
        return any;
    };
    const int e0a = im.pa_aabb[a];
    for (; p + 4 <= e0a; p += 4) {
        bool pass[4][K];
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; u++) any |= aabb_pass(p + u, pass[u]);
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int u = 0; u < 4; u++) enqueue(p + u, pass[u]);
        }
    }
    for (; p < e0a; p++) {
        bool pass[K];
        aabb_pass(p, pass);
        enqueue(p, pass);
    }
    for (; p + 2 <= e1; p += 2) {
        bool p0[K], p1[K];
        const bool any = box_pass(p, p0) | box_pass(p + 1, p1);
        if (__any_sync(0xffffffffu, any)) { enqueue(p, p0); enqueue(p + 1, p1); }
    }
    for (; p < e1; p++) {
        bool pass[K];
        box_pass(p, pass);
        enqueue(p, pass);
    }
    (void)e2; /* the moving partners were voted on by k_wave_fk */
    ws.publish(lane); /* the grid walk below appends lane by lane, through the shared counter */
#pragma unroll
    for (int k = 0; k < K; k++)
    if (G.r_on && live[k]) { /* one run of candidates: everything a geometry centred in this cell could touch */
        const Vec3<float> c = ca[k];
        const float fx = (c.x - G.r_ox) * G.r_inv_h, fy = (c.y - G.r_oy) * G.r_inv_h, fz = (c.z - G.r_oz) * G.r_inv_h;
        const int ix = (int)floorf(fx), iy = (int)floorf(fy), iz = (int)floorf(fz);
        if (ix >= 0 && iy >= 0 && iz >= 0 && ix < G.r_nx && iy < G.r_ny && iz < G.r_nz) {
            const int cell = (iz * G.r_ny + iy) * G.r_nx + ix;
            /* the query centre relative to the cell centre, like the entries */
            const float qx = c.x - (G.r_ox + ((float)ix + 0.5f) * G.r_h), qy = c.y - (G.r_oy + ((float)iy + 0.5f) * G.r_h),
                        qz = c.z - (G.r_oz + ((float)iz + 0.5f) * G.r_h);
            const uint32_t *live_row = G.live + (size_t)a * G.wps;
            const int k1 = G.r_start[cell + 1];
            for (int kk = G.r_start[cell]; kk < k1; kk++) {
                const uint2 e = G.r_c[kk];
                const float2 xy = __half22float2(*reinterpret_cast<const __half2 *>(&e.x));
                const float2 zr = __half22float2(*reinterpret_cast<const __half2 *>(&e.y));
                const float dx = xy.x - qx, dy = xy.y - qy, dz = zr.x - qz;
                const float reach = ra + zr.y;
                if (dx * dx + dy * dy + dz * dz > reach * reach) continue;
                const int b = G.r_id[kk]; /* rare from here on */
                if (!((live_row[b >> 5] >> (b & 31)) & 1u)) continue;
                ws.put(w, unc, atomicAdd(&cnt[warp], 1u), s[k], a, b, 1);
            }
        }
    }
    WaveStage::flush_cta(w, unc, (const uint2 *)smem, cnt, &s_base, tid, STAGE);
    (void)id_base;
}

/* quick certificates for the broadphase survivors */
__global__ void __launch_bounds__(256) k_wave_quick(Image<float> im, WaveBuf w, UncList unc, size_t id_base) {
    uint32_t n = *w.qcnt;
    if (n > w.queue_cap) n = w.queue_cap;
    const SQuick *SQ = (const SQuick *)im.sq;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 it = w.queue[i];
        const uint32_t s = it.x, key = it.y;
        if (w.hit[s]) continue;
        const int a = ITEM_A(key), b = ITEM_B(key);
        const DGeom<float> &A = im.mg[a];
        const int shape_b = ITEM_STATIC(key) ? im.sg[b].shape : im.mg[b].shape;
        int r = NEED;
        if (A.shape != K_MESH && shape_b != K_MESH) {
            const QG qa = qg_from_pose(A, wave_pose(w, a, s));
            if (ITEM_STATIC(key)) r = quick_pair(A.shape, qa, shape_b, qg_from_static(SQ[b]));
            else r = quick_pair(A.shape, qa, shape_b, qg_from_pose(im.mg[b], wave_pose(w, b, s)));
        }
        if (r == HIT) w.hit[s] = 1;
        else if (r == UNC) unc_emit(unc, (uint32_t)(id_base + s), 0u, key, TMK_WHOLE_PAIR);
        else if (r == NEED) {
            const auto g = cooperative_groups::coalesced_threads(); /* one atomic per warp, not per item */
            uint32_t k = 0;
            if (g.thread_rank() == 0) k = atomicAdd(w.pcnt, (uint32_t)g.size());
            k = g.shfl(k, 0) + g.thread_rank();
            if (k < w.pend_cap) w.pend[k] = it;
            else unc_emit(unc, (uint32_t)(id_base + s), 0u, key, TMK_WHOLE_PAIR);
        }
    }
}

/* records for the deferred narrowphase, only for states no certificate has condemned yet */
__global__ void __launch_bounds__(256) k_wave_export(Image<float> im, WaveBuf w, UncList unc, size_t id_base) {
    uint32_t n = *w.pcnt;
    if (n > w.pend_cap) n = w.pend_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 it = w.pend[i];
        const uint32_t s = it.x, key = it.y;
        if (w.hit[s]) continue;
        const int a = ITEM_A(key), b = ITEM_B(key);
        const int shape_b = ITEM_STATIC(key) ? im.sg[b].shape : im.mg[b].shape;
        const bool mesh = im.mg[a].shape == K_MESH || shape_b == K_MESH;
        uint32_t dst = 0;
        if (mesh) {
            const auto g = cooperative_groups::coalesced_threads();
            if (g.thread_rank() == 0) dst = atomicAdd(unc.mesh_n, (uint32_t)g.size());
            dst = g.shfl(dst, 0) + g.thread_rank();
        } else {
            const auto g = cooperative_groups::coalesced_threads();
            if (g.thread_rank() == 0) dst = atomicAdd(unc.conv_n, (uint32_t)g.size());
            dst = g.shfl(dst, 0) + g.thread_rank();
        }
        if (dst >= (mesh ? unc.mesh_cap : unc.conv_cap)) {
            unc_emit(unc, (uint32_t)(id_base + s), 0u, key, TMK_WHOLE_PAIR); /* list full: decide it exactly */
            continue;
        }
        const Xf<float> XA = wave_pose(w, a, s);
        Xf<float> XB = XA;
        if (!ITEM_STATIC(key)) XB = wave_pose(w, b, s);
        float4 *rec = (mesh ? unc.mesh_items : unc.conv_items) + (size_t)dst * TMK_MESH_REC;
        rec[0] = make_float4(__uint_as_float((uint32_t)(id_base + s)), __uint_as_float(0u), __uint_as_float(key), 0.f);
        rec[1] = make_float4(XA.r.x, XA.r.y, XA.r.z, XA.r.w);
        rec[2] = make_float4(XA.v.x, XA.v.y, XA.v.z, XB.v.x);
        rec[3] = make_float4(XB.r.x, XB.r.y, XB.r.z, XB.r.w);
        rec[4] = make_float4(XB.v.y, XB.v.z, 0.f, 0.f);
    }
}

/* validity words of the chunk: bit = no certificate says the state collides (the deferred kernels clear more) */
__global__ void __launch_bounds__(256) k_wave_bits(WaveBuf w, uint32_t n, uint32_t *bits_chunk) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < n && w.hit[s] == 0;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && s < n) bits_chunk[s >> 5] = m;
}

} /* namespace tmk */
