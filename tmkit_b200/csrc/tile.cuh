/* tile.cuh - the fused batched validity kernel (K1 + K2 + K3 per tile of configurations).
 *
 * One persistent CTA per resident slot walks tiles of TILE configurations:
 *   stage   scene tables (joints, moving / static geometry, pair list) copied ONCE per CTA into
 *           shared memory with cp.async.bulk (TMA 1-D bulk copies) completing on an mbarrier
 *   P1 FK   one thread per configuration: joint chain in registers, world poses of the moving
 *           geometries written to a shared slab [geom*7+k][thread] (bank-conflict free); they
 *           never touch HBM
 *   P2 broad  one thread per configuration, pair loop uniform across the warp (static bounds
 *           are shared-memory broadcasts); survivors of the bounding-sphere test are compacted
 *           with __ballot_sync/__popc into a shared work queue (one shared atomic per warp and
 *           pair with survivors)
 *   P3 narrow one thread per queued (configuration, pair): closed forms / SAT / GJK / BVH from
 *           dev_math.cuh on the staged geometry; queue order follows the class-sorted pair
 *           list, so warps mostly run one narrowphase class
 *   P4      one validity bit per configuration by ballot; undecided (UNC) configurations are
 *           appended to the list the double-precision kernel re-checks
 */
#pragma once
#include "kernels.cuh"

namespace tmk {

#define TMK_TILE 256
#define TMK_QCAP 8192
#define TMK_CHUNK 16

struct TileCfg {
    int off_joint, off_mg, off_sg, off_pairs, off_ssph, off_slab, off_queue, off_flags;
    int bytes_joint, bytes_mg, bytes_sg, bytes_pairs;
    int smem;
    int grid_max;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(TMK_TILE, 2)
k_validity_tile(Image<float> im, TileCfg cfg, QSrc qs, size_t n_cfg, uint32_t *bits, uint32_t *unc_list,
                uint32_t *unc_n) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t qcount;

    const int tid = threadIdx.x, lane = tid & 31;
    DJoint<float> *s_joint = (DJoint<float> *)(smem + cfg.off_joint);
    DGeom<float> *s_mg = (DGeom<float> *)(smem + cfg.off_mg);
    DGeom<float> *s_sg = (DGeom<float> *)(smem + cfg.off_sg);
    uint32_t *s_pairs = (uint32_t *)(smem + cfg.off_pairs);
    float4 *s_ssph = (float4 *)(smem + cfg.off_ssph);
    float *slab = (float *)(smem + cfg.off_slab);
    uint32_t *queue = (uint32_t *)(smem + cfg.off_queue);
    uint8_t *f_hit = smem + cfg.off_flags, *f_unc = f_hit + TMK_TILE;

    /* ---- stage the scene image: TMA bulk copies, one mbarrier ---- */
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t total = (uint32_t)(cfg.bytes_joint + cfg.bytes_mg + cfg.bytes_sg + cfg.bytes_pairs);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(total)
                     : "memory");
        bulk_g2s(s_joint, im.joints, (uint32_t)cfg.bytes_joint, &bar);
        bulk_g2s(s_mg, im.mg, (uint32_t)cfg.bytes_mg, &bar);
        if (cfg.bytes_sg) bulk_g2s(s_sg, im.sg, (uint32_t)cfg.bytes_sg, &bar);
        if (cfg.bytes_pairs) bulk_g2s(s_pairs, im.pairs, (uint32_t)cfg.bytes_pairs, &bar);
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0u)
                         : "memory");
        }
    }
    /* bounding spheres of the static geometry (world centre, radius) */
    for (int g = tid; g < im.n_sg; g += TMK_TILE) {
        const DGeom<float> &G = s_sg[g];
        Xf<float> X = xload(G.tf);
        Vec3<float> c = bound_centre(G, X);
        s_ssph[g] = make_float4(c.x, c.y, c.z, G.brad);
    }
    __syncthreads();

    Image<float> sm = im;
    sm.joints = s_joint; sm.mg = s_mg; sm.sg = s_sg; sm.pairs = s_pairs;

    const size_t n_tiles = (n_cfg + TMK_TILE - 1) / TMK_TILE;
    const float pad = Num<float>::eps() * 4.0f;

    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t c = tile * TMK_TILE + tid;
        const bool live = c < n_cfg && !im.static_hit;
        if (tid == 0) qcount = 0;
        f_hit[tid] = live ? 0 : 1;
        f_unc[tid] = 0;

        /* ---- P1: FK, poses to the slab ---- */
        if (live) {
            float q[TMK_MAXSUB];
            for (int k = 0; k < im.n_sub; k++) q[k] = load_q<float>(qs.p, qs.layout, qs.ld, c, k);
            Xf<float> jp[TMK_MAXJ];
            for (int j = 0; j < im.n_joint; j++) {
                const DJoint<float> &J = s_joint[j];
                Xf<float> E = xload(J.E);
                Xf<float> par;
                if (J.parent >= 0) par = jp[J.parent];
                jp[j] = fk_joint(E, mk<float>(J.ax[0], J.ax[1], J.ax[2]), J.off, J.type, q[J.qslot],
                                 J.parent >= 0 ? &par : (const Xf<float> *)nullptr);
            }
            for (int g = 0; g < im.n_mg; g++) {
                Xf<float> X = xmul(jp[s_mg[g].slot], xload(s_mg[g].tf));
                float *o = slab + (size_t)g * 7 * TMK_TILE + tid;
                o[0] = X.r.x; o[TMK_TILE] = X.r.y; o[2 * TMK_TILE] = X.r.z; o[3 * TMK_TILE] = X.r.w;
                o[4 * TMK_TILE] = X.v.x; o[5 * TMK_TILE] = X.v.y; o[6 * TMK_TILE] = X.v.z;
            }
        }
        __syncthreads();

        /* ---- P2 / P3 alternate over chunks of the pair list ---- */
        for (int p0 = 0; p0 < im.n_pair; p0 += TMK_CHUNK) {
            const int p1 = min(p0 + TMK_CHUNK, im.n_pair);
            const bool active = !f_hit[tid];
            for (int p = p0; p < p1; p++) {
                const uint32_t w = s_pairs[p];
                const int a = w & 0xff, b = (w >> 8) & 0xffff;
                bool pass = false;
                if (active) {
                    const float *pa = slab + (size_t)a * 7 * TMK_TILE + tid;
                    Vec3<float> ca = mk<float>(pa[4 * TMK_TILE], pa[5 * TMK_TILE], pa[6 * TMK_TILE]);
                    if (s_mg[a].shape == K_MESH) {
                        Quat<float> r;
                        r.x = pa[0]; r.y = pa[TMK_TILE]; r.z = pa[2 * TMK_TILE]; r.w = pa[3 * TMK_TILE];
                        ca = qrot(r, mk<float>(s_mg[a].dim[0], s_mg[a].dim[1], s_mg[a].dim[2])) + ca;
                    }
                    Vec3<float> cb;
                    float rb;
                    if ((w >> 24) & 1) {
                        float4 s = s_ssph[b];
                        cb = mk<float>(s.x, s.y, s.z);
                        rb = s.w;
                    } else {
                        const float *pb = slab + (size_t)b * 7 * TMK_TILE + tid;
                        cb = mk<float>(pb[4 * TMK_TILE], pb[5 * TMK_TILE], pb[6 * TMK_TILE]);
                        if (s_mg[b].shape == K_MESH) {
                            Quat<float> r;
                            r.x = pb[0]; r.y = pb[TMK_TILE]; r.z = pb[2 * TMK_TILE]; r.w = pb[3 * TMK_TILE];
                            cb = qrot(r, mk<float>(s_mg[b].dim[0], s_mg[b].dim[1], s_mg[b].dim[2])) + cb;
                        }
                        rb = s_mg[b].brad;
                    }
                    Vec3<float> d = ca - cb;
                    float reach = s_mg[a].brad + rb + pad;
                    pass = dot(d, d) <= reach * reach;
                }
                unsigned m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&qcount, (unsigned)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (pass) queue[base + __popc(m & ((1u << lane) - 1))] = ((uint32_t)tid << 20) | (uint32_t)p;
                }
            }
            __syncthreads();
            const uint32_t cnt = qcount;
            if (cnt + TMK_TILE * TMK_CHUNK > TMK_QCAP || p1 == im.n_pair) {
                /* ---- P3: narrowphase over the queue ---- */
                for (uint32_t i = tid; i < cnt; i += TMK_TILE) {
                    const uint32_t e = queue[i];
                    const int t = e >> 20, p = e & 0xfffff;
                    if (f_hit[t]) continue;
                    const uint32_t w = s_pairs[p];
                    const int a = w & 0xff, b = (w >> 8) & 0xffff;
                    const float *pa = slab + (size_t)a * 7 * TMK_TILE + t;
                    Xf<float> XA, XB;
                    XA.r.x = pa[0]; XA.r.y = pa[TMK_TILE]; XA.r.z = pa[2 * TMK_TILE]; XA.r.w = pa[3 * TMK_TILE];
                    XA.v.x = pa[4 * TMK_TILE]; XA.v.y = pa[5 * TMK_TILE]; XA.v.z = pa[6 * TMK_TILE];
                    const DGeom<float> *B;
                    if ((w >> 24) & 1) {
                        B = &s_sg[b];
                        XB = xload(B->tf);
                    } else {
                        B = &s_mg[b];
                        const float *pb = slab + (size_t)b * 7 * TMK_TILE + t;
                        XB.r.x = pb[0]; XB.r.y = pb[TMK_TILE]; XB.r.z = pb[2 * TMK_TILE]; XB.r.w = pb[3 * TMK_TILE];
                        XB.v.x = pb[4 * TMK_TILE]; XB.v.y = pb[5 * TMK_TILE]; XB.v.z = pb[6 * TMK_TILE];
                    }
                    int r = test_geoms(s_mg[a], XA, *B, XB, im.meshes);
                    if (r == HIT) f_hit[t] = 1;
                    else if (r == UNC) f_unc[t] = 1;
                }
                __syncthreads();
                if (tid == 0) qcount = 0;
                __syncthreads();
            }
        }

        /* ---- P4: verdict bits + uncertain list ---- */
        const bool hit = f_hit[tid] != 0, unc = !hit && f_unc[tid] != 0;
        unsigned free_mask = __ballot_sync(0xffffffffu, !hit && !unc);
        unsigned unc_mask = __ballot_sync(0xffffffffu, unc);
        if (lane == 0 && (c >> 5) < ((n_cfg + 31) >> 5)) bits[c >> 5] = free_mask;
        if (unc_mask) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(unc_n, (unsigned)__popc(unc_mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (unc) unc_list[base + __popc(unc_mask & ((1u << lane) - 1))] = (uint32_t)c;
        }
        __syncthreads();
    }
}

static inline int align16(int x) { return (x + 15) & ~15; }

static void tile_configure(TileCfg &cfg, const Image<float> &im) {
    int off = 0;
    cfg.bytes_joint = align16(im.n_joint * (int)sizeof(DJoint<float>));
    cfg.bytes_mg = align16(im.n_mg * (int)sizeof(DGeom<float>));
    cfg.bytes_sg = align16(im.n_sg * (int)sizeof(DGeom<float>));
    cfg.bytes_pairs = align16(im.n_pair * 4);
    cfg.off_joint = off; off += cfg.bytes_joint;
    cfg.off_mg = off; off += cfg.bytes_mg;
    cfg.off_sg = off; off += cfg.bytes_sg;
    cfg.off_pairs = off; off += cfg.bytes_pairs;
    cfg.off_ssph = off; off += align16(im.n_sg * 16);
    cfg.off_slab = off; off += im.n_mg * 7 * TMK_TILE * 4;
    cfg.off_queue = off; off += TMK_QCAP * 4;
    cfg.off_flags = off; off += 2 * TMK_TILE;
    cfg.smem = align16(off);
    cfg.grid_max = 0;
    if (cfg.smem > 227 * 1024) { cfg.smem = -1; return; } /* caller falls back to k_check_cfg */
    /* the attribute is per function, not per context: always raise it to the device maximum so that
     * images of different sizes (several collision contexts in one process) can coexist */
    int dev = 0, sms = 148, per_sm = 1, optin = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k_validity_tile) != cudaSuccess) { cfg.smem = -1; return; }
    int max_dyn = optin - (int)fa.sharedSizeBytes;
    if (cfg.smem > max_dyn ||
        cudaFuncSetAttribute(k_validity_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) != cudaSuccess) {
        cfg.smem = -1;
        return;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_validity_tile, TMK_TILE, cfg.smem);
    if (per_sm < 1) per_sm = 1;
    cfg.grid_max = sms * per_sm;
}

static void tile_launch(const TileCfg &cfg, const Image<float> &im, const QSrc &qs, size_t n_cfg, uint32_t *bits,
                        uint32_t *unc_list, uint32_t *unc_n, cudaStream_t s) {
    size_t n_tiles = (n_cfg + TMK_TILE - 1) / TMK_TILE;
    unsigned grid = (unsigned)std::min<size_t>(n_tiles, (size_t)cfg.grid_max);
    k_validity_tile<<<grid, TMK_TILE, cfg.smem, s>>>(im, cfg, qs, n_cfg, bits, unc_list, unc_n);
}

} /* namespace tmk */
