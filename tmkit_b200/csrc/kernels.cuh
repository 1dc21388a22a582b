/* kernels.cuh - device tables shared by all kernels, and the kernels that are NOT the batch pipeline
 * (that one lives in pipeline.cuh):
 *   k_fk_frames           K1 standalone (aa_rx_sg_tf_batch[_dev]): one thread per configuration, FP32 chain over the
 *                         varied joints (fixed frames folded on the host), one product per requested frame
 *   k_fk_single64         aa_rx_sg_tf: one configuration, double
 *   k_pairs_tf64          aa_rx_cl_check / allow_config / hoisted static pairs: one thread per geometry
 *                         pair on caller-provided absolute transforms, double
 *   k_check_cfg<T, LIST>  one thread per configuration, FK + all pairs inline: the instrumented path
 *                         (statistics, :track-collisions pair flags), the fallback for images the tile
 *                         kernel cannot hold, and (T = double, LIST) its exact re-check
 *   k_check_edges<T>,     one warp per edge / one thread per listed edge: the instrumented swept-edge
 *   k_recheck_edges64     path
 */
#pragma once
#include "dev_math.cuh"

namespace tmk {

#define TMK_MAXJ 24    /* device joints (varied configuration variables) */
#define TMK_MAXMG 48   /* moving collision geometries */
#define TMK_MAXSUB 24  /* length of a sub-configuration */

template <typename T> struct DJoint {
    T E[7];      /* pose relative to the parent joint (or world, parent < 0), fixed frames folded in */
    T ax[3];
    T off;
    int parent;  /* joint slot or -1 */
    int type;    /* 1 revolute, 2 prismatic */
    int qslot;   /* index into the sub-configuration */
    int pad;     /* > 0: slot+1 in the joint-pose slab (joints whose children do not follow them directly) */
};

template <typename T> struct DGeom {
    T tf[7];   /* moving: pose relative to joint `slot`; static: world pose; frame table: pose relative to `frame` */
    T dim[3];  /* sphere r | box half extents | cylinder radius, half height | mesh: bounding-sphere centre (local) */
    T brad;    /* bounding radius about the shape centre */
    int shape;
    int slot;
    int frame;
    int mesh;
};

/* pair word: bits 0-7 moving geom a | 8-23 geom b | 24 b is static | 25-27 shape of b | 28-31 class */
__host__ __device__ inline uint32_t pair_pack(int a, int b, int b_static, int cls, int shape_b) {
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)b_static << 24) | ((uint32_t)shape_b << 25) | ((uint32_t)cls << 28);
}
enum { PC_CLOSED = 0, PC_SAT = 1, PC_GJK = 2, PC_MESH = 3 };

/* Uniform grid over the small static geometry of large scenes (hundreds of obstacles): every gridded
 * geometry sits in the one cell that contains its bounding centre, a query covers the cells within
 * (radius of the moving geometry + r_max) of its centre. */
struct StaticGrid {
    int on;
    int nx, ny, nz;
    float ox, oy, oz, inv_h, r_max;
    const int *cell_start;         /* [nx*ny*nz + 1] */
    const unsigned short *items;   /* static geometry indices, cell by cell */
    const float4 *c4;              /* bounding centre + radius of items[k], in the same (cell) order: the walk reads
                                    * this array front to back and touches nothing else until a sphere test passes */
    int n_items;
    /* reach lists (wave.cuh): a finer grid whose cell lists hold EVERY item a moving geometry centred anywhere in the
     * cell could touch (items within the largest moving bounding radius + their own of the cell's box), as
     * (centre relative to the cell centre, radius) in half precision, rounded conservatively: one contiguous
     * run per query instead of a walk over dozens of nearly empty cells */
    int r_on, r_nx, r_ny, r_nz;
    float r_ox, r_oy, r_oz, r_h, r_inv_h;
    const int *r_start;            /* [r_nx*r_ny*r_nz + 1] */
    const uint2 *r_c;              /* half2(dx, dy), half2(dz, radius) */
    const unsigned short *r_id;    /* static geometry index of every entry */
    const uint32_t *live;          /* [n_mg][wps] bit b: pair (a, static b) is a candidate (not allowed-away, reachable) */
    int wps;
};

template <typename T> struct Image {
    const DJoint<T> *joints;
    const DGeom<T> *mg;
    const DGeom<T> *sg;
    const uint32_t *pairs;
    const MeshRef<T> *meshes;
    const void *sq;      /* SQuick[n_sg]: world-frame quick tables of the static geometry (float image) */
    const int *jg_begin; /* [n_joint+1]: moving geometries of joint j are mg[jg_begin[j] .. jg_begin[j+1]) */
    const int *pa_begin; /* [3*n_mg+1]: pairs owned by moving geometry a, in three segments - static partners
                          * tested by bounding sphere [3a], static boxes [3a+1], moving partners [3a+2] .. [3a+3] */
    const void *ps4;     /* float4[n_tpair]: bounding centre + radius of the static partner, in t_pairs order */
    const uint32_t *t_pairs; /* the pipeline's pair list: `pairs` minus the pairs the grid covers (same array without a grid) */
    const uint32_t *t_keys;  /* [n_tpair] the queue key of every pair (ITEM_KEY layout), ready to store (wave.cuh) */
    const void *ab4;         /* float4[2 * n_tpair]: world-aligned boxes of the box segment - centre, half extents (wave.cuh) */
    const int *pa_aabb;      /* [n_mg] end of the axis-aligned part of geometry a's box segment (pairs sorted: aligned first) */
    int n_tpair;
    StaticGrid grid;
    int n_joint, n_mg, n_sg, n_pair;
    int n_sub;
    int static_hit; /* a non-allowed static-static pair collides: every configuration is invalid */
};

struct QSrc {
    const void *p;
    const void *p1; /* edges: end points */
    int layout;     /* TMK_Q_* */
    size_t ld;
};

template <typename T> __device__ __forceinline__ T load_q(const void *p, int layout, size_t ld, size_t c, int k) {
    if (layout == 0) return (T)((const double *)p)[c * ld + k];
    if (layout == 1) return (T)((const float *)p)[c * ld + k];
    return (T)((const float *)p)[(size_t)k * ld + c];
}

struct DevStats {
    unsigned long long n_states, n_broad_pass, n_narrow[4], n_uncertain, gjk_iters;
};

template <typename T> __device__ __forceinline__ Shape<T> shape_of(const DGeom<T> &g) {
    Shape<T> s;
    s.kind = g.shape;
    s.a = g.dim[0]; s.b = g.dim[1]; s.c = g.dim[2];
    return s;
}

template <typename T> __device__ __forceinline__ Vec3<T> bound_centre(const DGeom<T> &g, const Xf<T> &X) {
    if (g.shape == K_MESH) return qrot(X.r, mk<T>(g.dim[0], g.dim[1], g.dim[2])) + X.v;
    return X.v;
}

/* narrowphase of two geometries with world poses */
template <typename T, class UncTri = NoTri, class NeedTri = NoTri>
__device__ int test_geoms(const DGeom<T> &A, const Xf<T> &XA, const DGeom<T> &B, const Xf<T> &XB,
                          const MeshRef<T> *meshes, UncTri on_unc_tri = UncTri(), NeedTri on_need_tri = NeedTri()) {
    if (A.shape != K_MESH && B.shape != K_MESH) return test_convex(shape_of(A), XA, shape_of(B), XB);
    if (A.shape == K_MESH && B.shape == K_MESH) {
        if (!Num<T>::exact) return UNC;
        /* exact path: every triangle of A against B's tree */
        const MeshRef<T> &MA = meshes[A.mesh], &MB = meshes[B.mesh];
        Xf<T> a_in_b = xrel(XB, XA);
        Mat3<T> R = q2m(a_in_b.r);
        int stackA[32], spA = 0;
        stackA[spA++] = 0;
        while (spA > 0) {
            BvhNode na = MA.nodes[stackA[--spA]];
            if (na.b <= 0) { stackA[spA++] = na.a; stackA[spA++] = -na.b - 1; continue; }
            for (int k = 0; k < na.b; k++) {
                Vec3<T> ta0, ta1, ta2;
                load_tri(MA.tris, na.a + k, ta0, ta1, ta2);
                Vec3<T> a0 = mv(R, ta0) + a_in_b.v;
                Vec3<T> a1 = mv(R, ta1) + a_in_b.v;
                Vec3<T> a2 = mv(R, ta2) + a_in_b.v;
                Vec3<T> c = (a0 + a1 + a2) * T(1.0 / 3.0);
                T r2 = t_max(dot(a0 - c, a0 - c), t_max(dot(a1 - c, a1 - c), dot(a2 - c, a2 - c)));
                T rad = t_sqrt(r2);
                int stackB[32], spB = 0;
                stackB[spB++] = 0;
                while (spB > 0) {
                    BvhNode nb = MB.nodes[stackB[--spB]];
                    T dx = t_max(t_max(T(nb.lo[0]) - c.x, c.x - T(nb.hi[0])), T(0));
                    T dy = t_max(t_max(T(nb.lo[1]) - c.y, c.y - T(nb.hi[1])), T(0));
                    T dz = t_max(t_max(T(nb.lo[2]) - c.z, c.z - T(nb.hi[2])), T(0));
                    if (dx * dx + dy * dy + dz * dz > rad * rad) continue;
                    if (nb.b <= 0) { stackB[spB++] = nb.a; stackB[spB++] = -nb.b - 1; continue; }
                    for (int j = 0; j < nb.b; j++) {
                        Vec3<T> b0, b1, b2;
                        load_tri(MB.tris, nb.a + j, b0, b1, b2);
                        if (test_tri_tri(a0, a1, a2, b0, b1, b2) == HIT) return HIT;
                    }
                }
            }
        }
        return SEP;
    }
    if (A.shape == K_MESH)
        return test_mesh_shape(meshes[A.mesh], XA, shape_of(B), XB, B.brad, (unsigned *)nullptr, on_unc_tri, on_need_tri);
    return test_mesh_shape(meshes[B.mesh], XB, shape_of(A), XA, A.brad, (unsigned *)nullptr, on_unc_tri, on_need_tri);
}

/* FK of the device joints + world poses of the moving geometries (thread-private arrays) */
template <typename T>
__device__ __forceinline__ void fk_config(const Image<T> &im, const T *q, Xf<T> *gp) {
    Xf<T> jp[TMK_MAXJ];
    for (int j = 0; j < im.n_joint; j++) {
        const DJoint<T> &J = im.joints[j];
        Xf<T> E = xload(J.E);
        Xf<T> par;
        if (J.parent >= 0) par = jp[J.parent];
        jp[j] = fk_joint(E, mk<T>(J.ax[0], J.ax[1], J.ax[2]), J.off, J.type, q[J.qslot], J.parent >= 0 ? &par : nullptr);
    }
    for (int g = 0; g < im.n_mg; g++) gp[g] = xmul(jp[im.mg[g].slot], xload(im.mg[g].tf));
}

/* world poses of two moving geometries only (the exact re-check of one pair): the chain is walked as far as the
 * later of their joints - parents precede children in the joint table - and no other geometry pose is built */
template <typename T>
__device__ __forceinline__ void fk_pair(const Image<T> &im, const T *q, int a, int b, Xf<T> &XA, Xf<T> &XB) {
    Xf<T> jp[TMK_MAXJ]; /* only joints with a child that does not follow them directly land here (local memory) */
    const int sa = im.mg[a].slot, sb = b >= 0 ? im.mg[b].slot : -1;
    const int last = sa > sb ? sa : sb;
    Xf<T> cur;
    cur.r.x = cur.r.y = cur.r.z = T(0); cur.r.w = T(1); cur.v = mk<T>(T(0), T(0), T(0));
    for (int j = 0; j <= last; j++) {
        const DJoint<T> &J = im.joints[j];
        Xf<T> par = cur;
        if (J.parent >= 0 && J.parent != j - 1) par = jp[J.parent];
        cur = fk_joint(xload(J.E), mk<T>(J.ax[0], J.ax[1], J.ax[2]), J.off, J.type, q[J.qslot], J.parent >= 0 ? &par : nullptr);
        if (J.pad) jp[j] = cur;
        if (j == sa) XA = xmul(cur, xload(im.mg[a].tf));
        if (j == sb) XB = xmul(cur, xload(im.mg[b].tf));
    }
}

/* validity of one configuration: SEP (free), HIT, UNC.  pairhit != NULL: no early-out, colliding
 * pairs are flagged (the :track-collisions feed-back). */
template <typename T>
__device__ int eval_config(const Image<T> &im, const T *q, uint8_t *pairhit, DevStats *st) {
    Xf<T> gp[TMK_MAXMG];
    fk_config(im, q, gp);
    int result = SEP;
    const T pad = Num<T>::eps() * T(4);
    for (int p = 0; p < im.n_pair; p++) {
        uint32_t w = im.pairs[p];
        int a = w & 0xff, b = (w >> 8) & 0xffff;
        const DGeom<T> &A = im.mg[a];
        const Xf<T> XA = gp[a];
        const DGeom<T> &B = (w >> 24) & 1 ? im.sg[b] : im.mg[b];
        const Xf<T> XB = (w >> 24) & 1 ? xload(B.tf) : gp[b];
        Vec3<T> d = bound_centre(A, XA) - bound_centre(B, XB);
        T reach = A.brad + B.brad + pad;
        if (dot(d, d) > reach * reach) continue;
        if (st) atomicAdd(&st->n_narrow[w >> 28], 1ull);
        int r = test_geoms(A, XA, B, XB, im.meshes);
        if (r == HIT) {
            if (!pairhit) return HIT;
            pairhit[p] = 1;
            result = HIT;
        } else if (r == UNC && result == SEP) result = UNC;
    }
    return result;
}

/* ------------------------------------------------------------------ K: one thread per configuration */
template <typename T, bool LIST>
__global__ void __launch_bounds__(128) k_check_cfg(Image<T> im, QSrc qs, size_t n_cfg, const uint32_t *list,
                                                   const uint32_t *list_n, uint32_t *bits, uint32_t *unc_list,
                                                   uint32_t *unc_n, uint8_t *pairhit, DevStats *st) {
    if (LIST) {
        size_t n = *list_n;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            size_t c = list[i];
            T q[TMK_MAXSUB];
            for (int k = 0; k < im.n_sub; k++) q[k] = load_q<T>(qs.p, qs.layout, qs.ld, c, k);
            int r = im.static_hit ? HIT : eval_config(im, q, pairhit, st);
            if (st) atomicAdd(&st->n_uncertain, 1ull);
            if (r == SEP) atomicOr(&bits[c >> 5], 1u << (c & 31));
        }
        return;
    }
    size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int r = HIT;
    if (c < n_cfg && !im.static_hit) {
        T q[TMK_MAXSUB];
        for (int k = 0; k < im.n_sub; k++) q[k] = load_q<T>(qs.p, qs.layout, qs.ld, c, k);
        r = eval_config(im, q, pairhit, st);
        if (st) atomicAdd(&st->n_states, 1ull);
    }
    unsigned free_mask = __ballot_sync(0xffffffffu, r == SEP);
    unsigned unc_mask = __ballot_sync(0xffffffffu, r == UNC);
    int lane = threadIdx.x & 31;
    if (lane == 0 && (c >> 5) < ((n_cfg + 31) >> 5)) bits[c >> 5] = free_mask;
    if (unc_mask) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(unc_n, (unsigned)__popc(unc_mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (r == UNC) unc_list[base + __popc(unc_mask & ((1u << lane) - 1))] = (uint32_t)c;
    }
}

/* ------------------------------------------------------------------ generic pair kernel on absolute transforms
 * (aa_rx_cl_check, allow_config, hoisted static pairs): one thread per geometry pair, double */
__global__ void k_pairs_tf64(const double *TF, const DGeom<double> *geoms, const uint2 *pairs, int n_pairs,
                             const MeshRef<double> *meshes, uint8_t *out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const DGeom<double> &A = geoms[pairs[p].x], &B = geoms[pairs[p].y];
    Xf<double> XA = xmul(xload(TF + 7 * (size_t)A.frame), xload(A.tf));
    Xf<double> XB = xmul(xload(TF + 7 * (size_t)B.frame), xload(B.tf));
    Vec3<double> d = bound_centre(A, XA) - bound_centre(B, XB);
    double reach = A.brad + B.brad;
    int r = SEP;
    if (dot(d, d) <= reach * reach) r = test_geoms(A, XA, B, XB, meshes);
    out[p] = (uint8_t)(r == HIT);
}

/* ------------------------------------------------------------------ FK over the full frame table (double) */
struct DFrame {
    double E[7];
    double ax[3];
    double off;
    int parent, type, qidx, pad;
};

/* single configuration: one thread walks the sorted frame table (the table is tiny and the
 * dependency chain is serial; this entry point exists for ABI compatibility, not throughput) */
__global__ void k_fk_single64(const DFrame *fr, int n_f, const double *q, double *rel, double *ab) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < n_f; i++) {
        DFrame F = fr[i];
        Xf<double> E = xload(F.E);
        Xf<double> r = F.type ? fk_joint(E, mk<double>(F.ax[0], F.ax[1], F.ax[2]), F.off, F.type, q[F.qidx],
                                          (const Xf<double> *)nullptr)
                              : E;
        Xf<double> a = r;
        if (F.parent >= 0) a = xmul(xload(ab + 7 * (size_t)F.parent), r);
        double *pr = rel + 7 * (size_t)i, *pa = ab + 7 * (size_t)i;
        pr[0] = r.r.x; pr[1] = r.r.y; pr[2] = r.r.z; pr[3] = r.r.w; pr[4] = r.v.x; pr[5] = r.v.y; pr[6] = r.v.z;
        pa[0] = a.r.x; pa[1] = a.r.y; pa[2] = a.r.z; pa[3] = a.r.w; pa[4] = a.v.x; pa[5] = a.v.y; pa[6] = a.v.z;
    }
}

/* K1 standalone (aa_rx_sg_tf_batch, aa_rx_sg_tf_batch_dev): one thread = one configuration.  The host folds every
 * fixed frame into the varied joint it hangs on (exactly as the collision image does), so the device work is the
 * chain of the <= TMK_MAXJ varied joints plus ONE transform product per requested frame: pose(frame) =
 * pose(joint slot) * fold(frame); frames no varied joint moves are constants.  Joint values come through load_q
 * (SoA float: one coalesced 128-byte transaction per joint and warp).  Output either as component planes
 * [frame * 7 + k][configuration] in float (coalesced: the layout to keep on the device) or as the dense double rows
 * [configuration][frame][7] of the host ABI. */
struct FkOut {
    float tf[7]; /* slot >= 0: pose relative to that joint | slot < 0: the constant world pose */
    int slot;
    int out;     /* position in the caller's frame list (the table is sorted by slot: a joint's frames are written while
                  * its pose is still in registers) */
};
#define TMK_TF_ROWS_F64 0
#define TMK_TF_SOA_F32 1

template <int TF_LAYOUT>
__global__ void __launch_bounds__(256) k_fk_frames(const DJoint<float> *joints, int n_joint, const FkOut *outs, int n_out, const void *Q,
                                                   int layout, size_t ld, size_t n_cfg, void *TF, size_t ld_tf) {
    extern __shared__ __align__(16) unsigned char fk_smem[];
    DJoint<float> *J = (DJoint<float> *)fk_smem;
    FkOut *O = (FkOut *)(fk_smem + (((size_t)n_joint * sizeof(DJoint<float>) + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < n_joint * (int)(sizeof(DJoint<float>) / 4); i += blockDim.x) ((uint32_t *)J)[i] = ((const uint32_t *)joints)[i];
    for (int i = threadIdx.x; i < n_out * (int)(sizeof(FkOut) / 4); i += blockDim.x) ((uint32_t *)O)[i] = ((const uint32_t *)outs)[i];
    __syncthreads();
    const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (c >= n_cfg) return;
    auto store = [&](int out, const Xf<float> &X) {
        if (TF_LAYOUT == TMK_TF_SOA_F32) {
            float *p = (float *)TF + (size_t)out * 7 * ld_tf + c;
            p[0] = X.r.x; p[ld_tf] = X.r.y; p[2 * ld_tf] = X.r.z; p[3 * ld_tf] = X.r.w;
            p[4 * ld_tf] = X.v.x; p[5 * ld_tf] = X.v.y; p[6 * ld_tf] = X.v.z;
        } else {
            double *p = (double *)TF + (c * (size_t)n_out + (size_t)out) * 7;
            p[0] = X.r.x; p[1] = X.r.y; p[2] = X.r.z; p[3] = X.r.w; p[4] = X.v.x; p[5] = X.v.y; p[6] = X.v.z;
        }
    };
    int k = 0;
    for (; k < n_out && O[k].slot < 0; k++) store(O[k].out, xload(O[k].tf)); /* frames no varied joint moves */
    /* The chain runs in registers: a joint's pose is the parent of the next one in a serial chain, and its frames are
     * written at once.  Only joints with a child that does not follow them directly (pad != 0) are kept in the
     * per-thread array - a dynamically indexed array lives in local memory, and with every joint in it the kernel moved
     * more local-memory bytes than poses (round 2: 54 % -> see profiles/README.md). */
    Xf<float> jp[TMK_MAXJ];
    Xf<float> cur;
    cur.r.x = cur.r.y = cur.r.z = 0.f; cur.r.w = 1.f; cur.v = mk<float>(0.f, 0.f, 0.f);
    for (int j = 0; j < n_joint; j++) {
        const DJoint<float> &Jt = J[j];
        Xf<float> par = cur;
        if (Jt.parent >= 0 && Jt.parent != j - 1) par = jp[Jt.parent];
        cur = fk_joint(xload(Jt.E), mk<float>(Jt.ax[0], Jt.ax[1], Jt.ax[2]), Jt.off, Jt.type, load_q<float>(Q, layout, ld, c, Jt.qslot),
                       Jt.parent >= 0 ? &par : (const Xf<float> *)nullptr);
        if (Jt.pad) jp[j] = cur;
        for (; k < n_out && O[k].slot == j; k++) store(O[k].out, xmul(cur, xload(O[k].tf)));
    }
}

#define TMK_ND_MAX (1 << 24) /* segments per edge are clamped here: 1.7e7 states on one edge is misuse, not a workload */
/* number of segments: ceil(|b-a|_2 / seg_len), at least 1 - evaluated in double with
 * round-to-nearest ops in index order so that it matches the host restatement bit for bit */
__device__ __forceinline__ int segment_count(const double *a, const double *b, int n, double seg_len) {
    double s = 0.0;
    for (int k = 0; k < n; k++) {
        double d = __dsub_rn(b[k], a[k]);
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
    const double r = ceil(__ddiv_rn(__dsqrt_rn(s), seg_len));
    if (!(r >= 1.0)) return 1;                      /* zero-length edge (or NaN end points: one state, Q1) */
    if (r > (double)TMK_ND_MAX) return TMK_ND_MAX;  /* the cast below is only defined inside int's range */
    return (int)r;
}

/* ------------------------------------------------------------------ K4: swept edges, one warp per edge */
template <typename T>
__global__ void __launch_bounds__(256) k_check_edges(Image<T> im, QSrc qs, size_t n_edge, double seg_len,
                                                     uint32_t *bits, int32_t *first_invalid, uint32_t *unc_list,
                                                     uint32_t *unc_n, DevStats *st) {
    __shared__ uint32_t word;
    if (threadIdx.x == 0) word = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int slot = warp; slot < 32; slot += nwarp) {
        size_t e = blockIdx.x * (size_t)32 + slot;
        if (e >= n_edge) break;
        double a[TMK_MAXSUB], b[TMK_MAXSUB];
        for (int k = 0; k < im.n_sub; k++) {
            a[k] = load_q<double>(qs.p, qs.layout, qs.ld, e, k);
            b[k] = load_q<double>(qs.p1, qs.layout, qs.ld, e, k);
        }
        int nd = segment_count(a, b, im.n_sub, seg_len);
        bool unc_before = false, decided = false;
        int first = -1;
        if (im.static_hit) { decided = true; first = 0; }
        for (int base = 0; base < nd && !decided; base += 32) {
            int i = base + lane; /* position in OMPL's test order: 0 = end state */
            int r = SEP;
            if (i < nd) {
                double t = 1.0;
                if (i > 0) t = __ddiv_rn((double)ompl_mid(nd, i - 1), (double)nd);
                T q[TMK_MAXSUB];
                for (int k = 0; k < im.n_sub; k++)
                    q[k] = (T)(i == 0 ? b[k] : __dadd_rn(a[k], __dmul_rn(__dsub_rn(b[k], a[k]), t)));
                r = eval_config(im, q, (uint8_t *)nullptr, st);
                if (st) atomicAdd(&st->n_states, 1ull);
            }
            unsigned hit = __ballot_sync(0xffffffffu, r == HIT);
            unsigned unc = __ballot_sync(0xffffffffu, r == UNC);
            if (hit) {
                int fl = __ffs(hit) - 1;
                first = base + fl;
                if (unc & ((1u << fl) - 1)) unc_before = true;
                decided = true;
            } else if (unc) unc_before = true;
        }
        if (lane == 0) {
            bool valid = !decided && !unc_before;
            /* uncertain: no certain hit but some state undecided, or (first index wanted and an
             * undecided state precedes the first certain hit) */
            bool recheck = (!decided && unc_before) || (decided && unc_before && first_invalid != nullptr);
            if (valid) atomicOr(&word, 1u << slot);
            if (first_invalid) first_invalid[e] = decided ? first : -1;
            if (recheck) unc_list[atomicAdd(unc_n, 1u)] = (uint32_t)e;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) bits[blockIdx.x] = word;
}

/* exact re-check of listed edges: one thread per edge, sequential in OMPL order with early-out */
__global__ void __launch_bounds__(128) k_recheck_edges64(Image<double> im, QSrc qs, double seg_len,
                                                         const uint32_t *list, const uint32_t *list_n,
                                                         uint32_t *bits, int32_t *first_invalid, DevStats *st) {
    size_t n = *list_n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t e = list[i];
        double a[TMK_MAXSUB], b[TMK_MAXSUB], q[TMK_MAXSUB];
        for (int k = 0; k < im.n_sub; k++) {
            a[k] = load_q<double>(qs.p, qs.layout, qs.ld, e, k);
            b[k] = load_q<double>(qs.p1, qs.layout, qs.ld, e, k);
        }
        int nd = segment_count(a, b, im.n_sub, seg_len);
        int first = -1;
        if (im.static_hit) first = 0;
        for (int t = 0; t < nd && first < 0; t++) {
            double s = 1.0;
            if (t > 0) s = __ddiv_rn((double)ompl_mid(nd, t - 1), (double)nd);
            for (int k = 0; k < im.n_sub; k++)
                q[k] = t == 0 ? b[k] : __dadd_rn(a[k], __dmul_rn(__dsub_rn(b[k], a[k]), s));
            if (eval_config(im, q, (uint8_t *)nullptr, st) == HIT) first = t;
            if (st) atomicAdd(&st->n_uncertain, 1ull);
        }
        if (first < 0) atomicOr(&bits[e >> 5], 1u << (e & 31));
        else atomicAnd(&bits[e >> 5], ~(1u << (e & 31)));
        if (first_invalid) first_invalid[e] = first;
    }
}

} /* namespace tmk */
