/* dev_math.cuh - device math of the validity check, templated on the scalar type.
 *
 *   T = float  : the fast path.  Every pair test returns SEP / HIT only when it holds a
 *                numerical CERTIFICATE that survives the FP32 error budget; otherwise UNC,
 *                and the configuration is re-evaluated by the T = double instantiation.
 *   T = double : exact verdicts (SEP / HIT), used for the single-configuration ABI, for the
 *                hoisted static pairs and for the uncertain configurations of a batch.
 *
 * Certificates (FP32):
 *   SEP  - a unit axis n whose support gap between the ORIGINAL shapes, evaluated directly
 *          along n, exceeds eps_free.  Valid whatever the quality of n.
 *   HIT  - closed forms: signed separation below -eps.  GJK: a point of the Minkowski
 *          difference of the ERODED shapes (a convex combination of support points,
 *          lambda >= 0) within hit_v of the origin, i.e. the originals interpenetrate by
 *          at least 2*erode - hit_v - error.  SAT: every axis overlaps by more than eps plus
 *          the direction error of that axis.
 *
 * The file compiles for the host too (g++, TMK_HD empty) so that tests can drive the very
 * same source against the CPU oracle without a GPU; that build is test infrastructure and
 * is never part of libtmkcl.so.
 */
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TMK_HD __host__ __device__ __forceinline__
#define TMK_HDN __host__ __device__ __noinline__
#else
#define TMK_HD inline
#define TMK_HDN
#endif

namespace tmk {

enum { SEP = 0, HIT = 1, UNC = 2 };
enum { K_SPHERE = 0, K_BOX = 1, K_CYL = 2, K_MESH = 3, K_TRI = 4 };

template <typename T> struct Num;
template <> struct Num<float> {
    static constexpr bool exact = false;
    static constexpr int gjk_maxit = 24;
    TMK_HD static float eps() { return 1.0e-5f; }      /* FP32 error allowance incl. the 1e-6 contract band */
    TMK_HD static float erode() { return 2.0e-5f; }    /* erosion of solids in GJK */
    TMK_HD static float hit_v() { return 5.0e-6f; }    /* |v| below this on eroded cores certifies a HIT */
    TMK_HD static float par2() { return 1.0e-6f; }     /* |L|^2 below this: cross axis too ill-conditioned */
    TMK_HD static float tiny() { return 1.0e-30f; }
    TMK_HD static float conv() { return 1.0e-6f; }
};
template <> struct Num<double> {
    static constexpr bool exact = true;
    static constexpr int gjk_maxit = 192;
    TMK_HD static double eps() { return 0.0; }
    TMK_HD static double erode() { return 0.0; }
    TMK_HD static double hit_v() { return 0.0; }
    TMK_HD static double par2() { return 1.0e-16; }
    TMK_HD static double tiny() { return 1.0e-30; }
    TMK_HD static double conv() { return 1.0e-13; }
};

TMK_HD float t_sqrt(float x) { return sqrtf(x); }
TMK_HD double t_sqrt(double x) { return sqrt(x); }
TMK_HD float t_abs(float x) { return fabsf(x); }
TMK_HD double t_abs(double x) { return fabs(x); }
TMK_HD float t_max(float a, float b) { return fmaxf(a, b); }
TMK_HD double t_max(double a, double b) { return fmax(a, b); }
TMK_HD float t_min(float a, float b) { return fminf(a, b); }
TMK_HD double t_min(double a, double b) { return fmin(a, b); }
/* sin / cos of a joint half-angle: Cody-Waite reduction by pi/2 and the classic minimax polynomials on
 * [-pi/4, pi/4] - about 1e-7 absolute for |a| < 100, a third of the instructions of sincosf (whose
 * large-argument path is never needed here).  Larger arguments fall back to sincosf. */
TMK_HD void t_sincos(float a, float *s, float *c) {
    if (!(fabsf(a) < 100.0f)) { sincosf(a, s, c); return; }
    const float k = rintf(a * 0.636619772367581343f);
    float r = fmaf(-k, 1.57079637050628662109375f, a);  /* float(pi/2) ... */
    r = fmaf(-k, -4.37113882867379e-8f, r);            /* ... and the rest of pi/2 */
    const float r2 = r * r;
    const float sn = fmaf(r * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
    const float cs = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                          fmaf(-0.5f, r2, 1.0f));
    const int q = (int)k;
    const float s0 = (q & 1) ? cs : sn, c0 = (q & 1) ? sn : cs;
    *s = (q & 2) ? -s0 : s0;
    *c = ((q + 1) & 2) ? -c0 : c0;
}
TMK_HD void t_sincos(double a, double *s, double *c) { sincos(a, s, c); }

/* ------------------------------------------------------------------ vectors, quaternions, poses */
template <typename T> struct Vec3 { T x, y, z; };
template <typename T> TMK_HD Vec3<T> mk(T x, T y, T z) { Vec3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> TMK_HD Vec3<T> operator+(Vec3<T> a, Vec3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> TMK_HD Vec3<T> operator-(Vec3<T> a, Vec3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> TMK_HD Vec3<T> operator-(Vec3<T> a) { return mk<T>(-a.x, -a.y, -a.z); }
template <typename T> TMK_HD Vec3<T> operator*(Vec3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> TMK_HD T dot(Vec3<T> a, Vec3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> TMK_HD Vec3<T> cross(Vec3<T> a, Vec3<T> b) {
    return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

template <typename T> struct Quat { T x, y, z, w; };
template <typename T> struct Xf { Quat<T> r; Vec3<T> v; };

template <typename T> TMK_HD Quat<T> qmul(Quat<T> a, Quat<T> b) {
    Quat<T> c;
    c.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    c.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    c.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    c.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return c;
}
template <typename T> TMK_HD Vec3<T> qrot(Quat<T> q, Vec3<T> v) {
    Vec3<T> u = mk<T>(q.x, q.y, q.z);
    Vec3<T> t = cross(u, v) * T(2);
    return v + t * q.w + cross(u, t);
}
template <typename T> TMK_HD Quat<T> qconj(Quat<T> q) { Quat<T> c; c.x = -q.x; c.y = -q.y; c.z = -q.z; c.w = q.w; return c; }
/* a * b : first b, then a   */
template <typename T> TMK_HD Xf<T> xmul(const Xf<T> &a, const Xf<T> &b) {
    Xf<T> c;
    c.r = qmul(a.r, b.r);
    c.v = qrot(a.r, b.v) + a.v;
    return c;
}
/* inverse(a) * b */
template <typename T> TMK_HD Xf<T> xrel(const Xf<T> &a, const Xf<T> &b) {
    Xf<T> c;
    Quat<T> ai = qconj(a.r);
    c.r = qmul(ai, b.r);
    c.v = qrot(ai, b.v - a.v);
    return c;
}
template <typename T> TMK_HD Xf<T> xload(const T *p) {
    Xf<T> x;
    x.r.x = p[0]; x.r.y = p[1]; x.r.z = p[2]; x.r.w = p[3];
    x.v.x = p[4]; x.v.y = p[5]; x.v.z = p[6];
    return x;
}

template <typename T> struct Mat3 { T m00, m01, m02, m10, m11, m12, m20, m21, m22; };
template <typename T> TMK_HD Mat3<T> q2m(Quat<T> q) {
    T n = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    T s = T(2) / n;
    Mat3<T> R;
    R.m00 = 1 - s * (q.y * q.y + q.z * q.z); R.m01 = s * (q.x * q.y - q.z * q.w); R.m02 = s * (q.x * q.z + q.y * q.w);
    R.m10 = s * (q.x * q.y + q.z * q.w); R.m11 = 1 - s * (q.x * q.x + q.z * q.z); R.m12 = s * (q.y * q.z - q.x * q.w);
    R.m20 = s * (q.x * q.z - q.y * q.w); R.m21 = s * (q.y * q.z + q.x * q.w); R.m22 = 1 - s * (q.x * q.x + q.y * q.y);
    return R;
}
template <typename T> TMK_HD Vec3<T> mv(const Mat3<T> &R, Vec3<T> v) {
    return mk<T>(R.m00 * v.x + R.m01 * v.y + R.m02 * v.z, R.m10 * v.x + R.m11 * v.y + R.m12 * v.z,
                 R.m20 * v.x + R.m21 * v.y + R.m22 * v.z);
}
template <typename T> TMK_HD Vec3<T> mtv(const Mat3<T> &R, Vec3<T> v) {
    return mk<T>(R.m00 * v.x + R.m10 * v.y + R.m20 * v.z, R.m01 * v.x + R.m11 * v.y + R.m21 * v.z,
                 R.m02 * v.x + R.m12 * v.y + R.m22 * v.z);
}

/* ------------------------------------------------------------------ forward kinematics of one joint
 * rel = E * Rot(axis, q+offset)  |  E * Trans(axis*(q+offset));  abs = parent * rel   (A.2) */
template <typename T>
TMK_HD Xf<T> fk_joint(const Xf<T> &E, Vec3<T> axis, T offset, int type, T q, const Xf<T> *parent) {
    Xf<T> rel;
    T a = q + offset;
    if (type == 1) {
        T s, c;
        t_sincos(T(0.5) * a, &s, &c);
        if (axis.x == T(0) && axis.y == T(0) && axis.z == T(1)) { /* URDF-style joints: half the multiplies */
            rel.r.x = E.r.x * c + E.r.y * s;
            rel.r.y = E.r.y * c - E.r.x * s;
            rel.r.z = E.r.z * c + E.r.w * s;
            rel.r.w = E.r.w * c - E.r.z * s;
        } else {
            Quat<T> qa;
            qa.x = axis.x * s; qa.y = axis.y * s; qa.z = axis.z * s; qa.w = c;
            rel.r = qmul(E.r, qa);
        }
        rel.v = E.v;
    } else {
        rel.r = E.r;
        rel.v = E.v + qrot(E.r, axis * a);
    }
    if (!parent) return rel;
    return xmul(*parent, rel);
}

/* ------------------------------------------------------------------ shapes */
template <typename T> struct Shape {
    int kind;  /* K_SPHERE: a = radius | K_BOX: a,b,c = half extents | K_CYL: a = radius, b = half height (centred) */
    T a, b, c;
};

template <typename T> TMK_HD int classify(T s) {
    if (Num<T>::exact) return s > T(0) ? SEP : HIT;
    if (s > Num<T>::eps()) return SEP;
    if (s < -Num<T>::eps()) return HIT;
    return UNC;
}

/* distance from point p (shape-local) to a solid; 0 inside */
template <typename T> TMK_HD T point_box(Vec3<T> p, T hx, T hy, T hz) {
    T ex = t_max(t_abs(p.x) - hx, T(0)), ey = t_max(t_abs(p.y) - hy, T(0)), ez = t_max(t_abs(p.z) - hz, T(0));
    return t_sqrt(ex * ex + ey * ey + ez * ez);
}
template <typename T> TMK_HD T point_cyl(Vec3<T> p, T r, T hh) {
    T er = t_max(t_sqrt(p.x * p.x + p.y * p.y) - r, T(0)), ez = t_max(t_abs(p.z) - hh, T(0));
    return t_sqrt(er * er + ez * ez);
}
/* closest point on triangle (a,b,c) to p - Voronoi regions */
template <typename T> TMK_HD Vec3<T> closest_on_tri(Vec3<T> p, Vec3<T> a, Vec3<T> b, Vec3<T> c) {
    Vec3<T> ab = b - a, ac = c - a, ap = p - a;
    T d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return a;
    Vec3<T> bp = p - b;
    T d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return b;
    T vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) return a + ab * (d1 / (d1 - d3));
    Vec3<T> cp = p - c;
    T d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return c;
    T vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) return a + ac * (d2 / (d2 - d6));
    T va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
    T den = T(1) / (va + vb + vc);
    return a + ab * (vb * den) + ac * (vc * den);
}

/* ------------------------------------------------------------------ box-box: 15-axis SAT in A's frame */
template <typename T>
TMK_HD int sat_box_box(T ax, T ay, T az, T bx, T by, T bz, const Mat3<T> &R, Vec3<T> t) {
    /* R: B's axes expressed in A's frame (columns), t: B's centre in A's frame */
    const T ha[3] = {ax, ay, az}, hb[3] = {bx, by, bz};
    const T Rm[3][3] = {{R.m00, R.m01, R.m02}, {R.m10, R.m11, R.m12}, {R.m20, R.m21, R.m22}};
    const T tv[3] = {t.x, t.y, t.z};
    T g = -INFINITY;  /* largest certified gap */
    bool all_overlap = true;
    const T eps = Num<T>::eps();
    const T scale = t_abs(t.x) + t_abs(t.y) + t_abs(t.z) + ax + ay + az + bx + by + bz;
    /* A's face axes */
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T rb = hb[0] * t_abs(Rm[i][0]) + hb[1] * t_abs(Rm[i][1]) + hb[2] * t_abs(Rm[i][2]);
        T s = t_abs(tv[i]) - ha[i] - rb;
        g = t_max(g, s);
        if (!(s < -eps)) all_overlap = false;
    }
    /* B's face axes */
#pragma unroll
    for (int j = 0; j < 3; j++) {
        T ra = ha[0] * t_abs(Rm[0][j]) + ha[1] * t_abs(Rm[1][j]) + ha[2] * t_abs(Rm[2][j]);
        T tp = tv[0] * Rm[0][j] + tv[1] * Rm[1][j] + tv[2] * Rm[2][j];
        T s = t_abs(tp) - ra - hb[j];
        g = t_max(g, s);
        if (!(s < -eps)) all_overlap = false;
    }
    if (g > eps) return SEP;
    /* edge-edge axes n = a_i x b_j, normalised explicitly, gap evaluated along n directly */
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
            /* a_i x b_j in A's frame: e_i x (R col j) */
            T n[3];
            n[i] = T(0);
            n[i1] = -Rm[i2][j];
            n[i2] = Rm[i1][j];
            T L2 = n[i1] * n[i1] + n[i2] * n[i2];
            if (L2 < Num<T>::par2()) { if (!Num<T>::exact) all_overlap = false; continue; }
            T inv = T(1) / t_sqrt(L2);
            n[i1] *= inv; n[i2] *= inv;
            T ra = ha[i1] * t_abs(n[i1]) + ha[i2] * t_abs(n[i2]);
            T rb = T(0);
#pragma unroll
            for (int k = 0; k < 3; k++) rb += hb[k] * t_abs(n[i1] * Rm[i1][k] + n[i2] * Rm[i2][k]);
            T s = t_abs(tv[i1] * n[i1] + tv[i2] * n[i2]) - ra - rb;
            g = t_max(g, s);
            /* direction error of n grows as 1/|L| in FP32 */
            T need = Num<T>::exact ? T(0) : eps + scale * T(2.0e-7) * inv;
            if (!(s < -need)) all_overlap = false;
        }
    }
    if (g > eps) return SEP;
    if (Num<T>::exact) return HIT;
    return all_overlap ? HIT : UNC;
}

/* ------------------------------------------------------------------ GJK in A's local frame
 * A: axis-aligned box or cylinder at the origin.  B: box / cylinder with pose (R,t) in A's
 * frame, or a triangle given by three points in A's frame. */
template <typename T> struct GjkB {
    int kind;
    T h0, h1, h2;
    Mat3<T> R;
    Vec3<T> t;
    Vec3<T> p0, p1, p2;
};

template <typename T> TMK_HD Vec3<T> sup_local(int kind, T h0, T h1, T h2, Vec3<T> d) {
    if (kind == K_BOX) return mk<T>(d.x >= 0 ? h0 : -h0, d.y >= 0 ? h1 : -h1, d.z >= 0 ? h2 : -h2);
    /* cylinder: h0 = radius, h1 = half height */
    T rho2 = d.x * d.x + d.y * d.y;
    T z = d.z >= 0 ? h1 : -h1;
    if (rho2 > T(0)) {
        T s = h0 / t_sqrt(rho2);
        return mk<T>(d.x * s, d.y * s, z);
    }
    return mk<T>(T(0), T(0), z);
}
template <typename T> TMK_HD Vec3<T> sup_B(const GjkB<T> &B, Vec3<T> d) {
    if (B.kind == K_TRI) {
        T a = dot(B.p0, d), b = dot(B.p1, d), c = dot(B.p2, d);
        if (a >= b && a >= c) return B.p0;
        return b >= c ? B.p1 : B.p2;
    }
    return mv(B.R, sup_local(B.kind, B.h0, B.h1, B.h2, mtv(B.R, d))) + B.t;
}

/* closest point of simplex W[0..n) to the origin; reduces W; returns new n; n==4 => origin inside,
 * lam[] then holds its barycentric coordinates */
template <typename T> TMK_HD int tri_region(Vec3<T> a, Vec3<T> b, Vec3<T> c, Vec3<T> &v, T lam[3]) {
    Vec3<T> ab = b - a, ac = c - a;
    T d1 = -dot(ab, a), d2 = -dot(ac, a);
    lam[0] = lam[1] = lam[2] = T(0);
    if (d1 <= 0 && d2 <= 0) { v = a; lam[0] = 1; return 1; }
    T d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0 && d4 <= d3) { v = b; lam[1] = 1; return 2; }
    T vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        T s = d1 / (d1 - d3);
        v = a + ab * s; lam[0] = 1 - s; lam[1] = s;
        return 3;
    }
    T d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0 && d5 <= d6) { v = c; lam[2] = 1; return 4; }
    T vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        T s = d2 / (d2 - d6);
        v = a + ac * s; lam[0] = 1 - s; lam[2] = s;
        return 5;
    }
    T va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        T s = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        v = b + (c - b) * s; lam[1] = 1 - s; lam[2] = s;
        return 6;
    }
    T den = T(1) / (va + vb + vc);
    T s = vb * den, u = vc * den;
    v = a + ab * s + ac * u;
    lam[0] = 1 - s - u; lam[1] = s; lam[2] = u;
    return 7;
}

template <typename T> TMK_HD int simplex_solve(Vec3<T> W[4], int n, Vec3<T> &v, bool &inside) {
    inside = false;
    if (n == 1) { v = W[0]; return 1; }
    if (n == 2) {
        Vec3<T> ab = W[1] - W[0];
        T t = -dot(W[0], ab), den = dot(ab, ab);
        if (t <= 0 || den <= 0) { v = W[0]; return 1; }
        if (t >= den) { v = W[1]; W[0] = W[1]; return 1; }
        v = W[0] + ab * (t / den);
        return 2;
    }
    T lam[3];
    if (n == 3) {
        int m = tri_region(W[0], W[1], W[2], v, lam);
        int k = 0;
        Vec3<T> t0 = W[0], t1 = W[1], t2 = W[2];
        if (m & 1) W[k++] = t0;
        if (m & 2) W[k++] = t1;
        if (m & 4) W[k++] = t2;
        return k;
    }
    /* tetrahedron: faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0) */
    T best = INFINITY;
    Vec3<T> bv = mk<T>(0, 0, 0);
    int bm = 0, bf = -1;
    bool any_out = false, degenerate = false;
    T sc = T(0);
#pragma unroll
    for (int k = 0; k < 4; k++) sc = t_max(sc, t_max(t_abs(W[k].x), t_max(t_abs(W[k].y), t_abs(W[k].z))));
    const T dtol = (Num<T>::exact ? T(1e-14) : T(1e-6)) * sc * sc * sc;
#pragma unroll
    for (int f = 0; f < 4; f++) {
        const int i0 = (f == 3) ? 1 : 0, i1 = (f == 0) ? 1 : (f == 1) ? 2 : 3, i2 = (f == 0) ? 2 : (f == 1) ? 3 : (f == 2) ? 1 : 2,
                  i3 = (f == 0) ? 3 : (f == 1) ? 1 : (f == 2) ? 2 : 0;
        Vec3<T> a = W[i0], b = W[i1], c = W[i2], d = W[i3];
        Vec3<T> nrm = cross(b - a, c - a);
        T sp = -dot(a, nrm), sd = dot(d - a, nrm);
        if (t_abs(sd) <= dtol) degenerate = true;
        if (degenerate || sp * sd < 0 || sp == T(0)) {
            Vec3<T> cv;
            int m = tri_region(a, b, c, cv, lam);
            T dd = dot(cv, cv);
            any_out = true;
            if (dd < best) { best = dd; bv = cv; bm = m; bf = f; }
        }
    }
    if (degenerate) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const int i0 = (f == 3) ? 1 : 0, i1 = (f == 0) ? 1 : (f == 1) ? 2 : 3,
                      i2 = (f == 0) ? 2 : (f == 1) ? 3 : (f == 2) ? 1 : 2;
            Vec3<T> cv;
            int m = tri_region(W[i0], W[i1], W[i2], cv, lam);
            T dd = dot(cv, cv);
            if (dd < best) { best = dd; bv = cv; bm = m; bf = f; }
        }
    }
    if (!any_out && !degenerate) { inside = true; v = mk<T>(0, 0, 0); return 4; }
    const int i0 = (bf == 3) ? 1 : 0, i1 = (bf == 0) ? 1 : (bf == 1) ? 2 : 3, i2 = (bf == 0) ? 2 : (bf == 1) ? 3 : (bf == 2) ? 1 : 2;
    Vec3<T> t0 = W[i0], t1 = W[i1], t2 = W[i2];
    int k = 0;
    if (bm & 1) W[k++] = t0;
    if (bm & 2) W[k++] = t1;
    if (bm & 4) W[k++] = t2;
    v = bv;
    return k;
}

/* point of tetrahedron W closest to the origin expressed through clamped barycentric
 * coordinates - the FP32 HIT certificate when the origin is reported inside */
template <typename T> TMK_HD T inside_residual2(const Vec3<T> W[4]) {
    Vec3<T> a = W[0] - W[3], b = W[1] - W[3], c = W[2] - W[3], r = -W[3];
    Vec3<T> bc = cross(b, c);
    T det = dot(a, bc);
    if (t_abs(det) < Num<T>::tiny()) return INFINITY;
    T l0 = dot(r, bc) / det, l1 = dot(a, cross(r, c)) / det, l2 = dot(a, cross(b, r)) / det;
    T l3 = T(1) - l0 - l1 - l2;
    l0 = t_max(l0, T(0)); l1 = t_max(l1, T(0)); l2 = t_max(l2, T(0)); l3 = t_max(l3, T(0));
    T s = T(1) / (l0 + l1 + l2 + l3);
    Vec3<T> p = (W[0] * l0 + W[1] * l1 + W[2] * l2 + W[3] * l3) * s;
    return dot(p, p);
}

/* kindA / (a0,a1,a2): A's ORIGINAL dims; B holds B's ORIGINAL dims.  n_solid = number of solids
 * (2, or 1 when B is a triangle). */
template <typename T> TMK_HD int gjk_pair(int kindA, T a0, T a1, T a2, GjkB<T> B, int *iters = nullptr) {
    const T er = Num<T>::erode();
    a0 -= er; a1 -= er; a2 -= er;
    T e_sum = T(1.75) * er;
    if (B.kind != K_TRI) { B.h0 -= er; B.h1 -= er; B.h2 -= er; e_sum += T(1.75) * er; }
    const T t_free = e_sum + Num<T>::eps();
    const T tf2 = t_free * t_free;
    const T hv2 = Num<T>::hit_v() * Num<T>::hit_v();

    Vec3<T> v = (B.kind == K_TRI) ? -B.p0 : -B.t; /* centre(A) - point(B) */
    if (dot(v, v) < T(1e-12)) v = mk<T>(T(1), T(0), T(0));
    Vec3<T> W[4];
    int n = 0;
    int result = UNC;
    int it = 0;
    for (; it < Num<T>::gjk_maxit; it++) {
        Vec3<T> w = sup_local(kindA, a0, a1, a2, -v) - sup_B(B, v);
        T vv = dot(v, v), vw = dot(v, w);
        if (vw > T(0) && vw * vw > tf2 * vv) { result = SEP; break; }
        if (n > 0) {
            if (Num<T>::exact) {
                if (vv <= Num<T>::tiny()) { result = HIT; break; }
            } else if (vv < hv2) { result = HIT; break; }
            if (vv - vw <= Num<T>::conv() * vv) {
                /* converged to the distance |v| with neither certificate */
                result = Num<T>::exact ? (vv > T(0) ? SEP : HIT) : UNC;
                break;
            }
        }
        bool dup = false;
        for (int k = 0; k < n; k++) {
            Vec3<T> d = w - W[k];
            if (dot(d, d) <= (Num<T>::exact ? T(1e-28) : T(1e-13)) * (T(1) + dot(w, w))) dup = true;
        }
        if (dup) { result = Num<T>::exact ? (vv > T(0) ? SEP : HIT) : UNC; break; }
        W[n++] = w;
        Vec3<T> nv;
        bool inside;
        Vec3<T> Wsave[4] = {W[0], W[1], W[2], W[3]};
        n = simplex_solve(W, n, nv, inside);
        if (inside) {
            if (Num<T>::exact) result = HIT;
            else result = inside_residual2(Wsave) < hv2 ? HIT : UNC;
            break;
        }
        T nn = dot(nv, nv);
        if (it > 0 && nn >= vv) { result = Num<T>::exact ? (vv > T(0) ? SEP : HIT) : UNC; break; }
        v = nv;
    }
    if (iters) *iters = it + 1;
    if (it == Num<T>::gjk_maxit && Num<T>::exact) result = dot(v, v) > T(0) ? SEP : HIT;
    return result;
}

/* ------------------------------------------------------------------ convex pair dispatch (world poses) */
template <typename T> struct NarrowStats { unsigned n_closed, n_sat, n_gjk, n_leaf, gjk_iters; };

template <typename T>
TMK_HD int test_tri_vs_shape(const Shape<T> &S, Vec3<T> p0, Vec3<T> p1, Vec3<T> p2) {
    /* triangle points already in S's local frame */
    if (S.kind == K_SPHERE) {
        Vec3<T> c = closest_on_tri(mk<T>(0, 0, 0), p0, p1, p2);
        return classify(t_sqrt(dot(c, c)) - S.a);
    }
    GjkB<T> B;
    B.kind = K_TRI;
    B.p0 = p0; B.p1 = p1; B.p2 = p2;
    B.h0 = B.h1 = B.h2 = T(0);
    return gjk_pair<T>(S.kind, S.a, S.kind == K_BOX ? S.b : S.b, S.kind == K_BOX ? S.c : T(0), B);
}

template <typename T>
TMK_HD int test_convex(const Shape<T> &A, const Xf<T> &XA, const Shape<T> &B, const Xf<T> &XB, int *iters = nullptr) {
    /* order so that kind(A) <= kind(B) */
    if (A.kind > B.kind) return test_convex(B, XB, A, XA, iters);
    if (A.kind == K_SPHERE) {
        if (B.kind == K_SPHERE) {
            Vec3<T> d = XA.v - XB.v;
            return classify(t_sqrt(dot(d, d)) - A.a - B.a);
        }
        Vec3<T> p = qrot(qconj(XB.r), XA.v - XB.v); /* sphere centre in B's frame */
        T d = (B.kind == K_BOX) ? point_box(p, B.a, B.b, B.c) : point_cyl(p, B.a, B.b);
        return classify(d - A.a);
    }
    Xf<T> rel = xrel(XA, XB); /* B in A's frame */
    Mat3<T> R = q2m(rel.r);
    if (A.kind == K_BOX && B.kind == K_BOX) return sat_box_box(A.a, A.b, A.c, B.a, B.b, B.c, R, rel.v);
    GjkB<T> G;
    G.kind = B.kind;
    G.h0 = B.a; G.h1 = B.b; G.h2 = (B.kind == K_BOX) ? B.c : T(0);
    G.R = R;
    G.t = rel.v;
    G.p0 = G.p1 = G.p2 = mk<T>(0, 0, 0);
    return gjk_pair<T>(A.kind, A.a, A.b, (A.kind == K_BOX) ? A.c : T(0), G, iters);
}

/* ------------------------------------------------------------------ quick certificates (FP32 broadphase stage)
 * World-frame quick geometry: centre c (bounding centre for meshes), rotation columns x,y,z,
 * dims (sphere: h0 = r | box: half extents | cylinder: h0 = r, h1 = half height), bounding radius. */
#define NEED 3 /* quick test: undecided, run the narrowphase */
struct QG {
    Vec3<float> c, x, y, z;
    float h0, h1, h2, brad;
};

TMK_HD QG qg_make(int kind, float h0, float h1, float h2, float brad, const Xf<float> &X, Vec3<float> mesh_centre) {
    QG g;
    Mat3<float> R = q2m(X.r);
    g.x = mk<float>(R.m00, R.m10, R.m20);
    g.y = mk<float>(R.m01, R.m11, R.m21);
    g.z = mk<float>(R.m02, R.m12, R.m22);
    g.c = X.v;
    if (kind == K_MESH) g.c = mv(R, mesh_centre) + X.v;
    g.h0 = h0; g.h1 = h1; g.h2 = h2;
    g.brad = brad;
    return g;
}

/* support radius of a cylinder (axis u, radius r, half height hh) along the unit vector n */
TMK_HD float cyl_proj(Vec3<float> u, float r, float hh, Vec3<float> n) {
    Vec3<float> c = cross(n, u);
    return hh * fabsf(dot(u, n)) + r * sqrtf(dot(c, c));
}
TMK_HD float box_proj(const QG &B, Vec3<float> n) {
    return B.h0 * fabsf(dot(B.x, n)) + B.h1 * fabsf(dot(B.y, n)) + B.h2 * fabsf(dot(B.z, n));
}

/* Quick verdict of one pair from world-frame quick geometry; ka <= kb is NOT required.
 * SEP / HIT carry the same certificates as dev_math.cuh (a unit axis with gap > eps; a point of one
 * shape inside the other by more than eps; closed forms beyond eps).  UNC: a closed form inside
 * the FP32 allowance.  NEED: undecided, run the full narrowphase. */
TMK_HD int quick_pair(int ka_in, const QG &A_in, int kb_in, const QG &B_in) {
    const float eps = Num<float>::eps();
    {
        Vec3<float> d0 = B_in.c - A_in.c;
        float reach = A_in.brad + B_in.brad + eps * 4.0f;
        if (dot(d0, d0) > reach * reach) return SEP;
    }
    if (ka_in == K_MESH || kb_in == K_MESH) return NEED;
    /* order so that kind(A) <= kind(B) (the pair's shapes are warp-uniform, so is this choice) */
    const bool sw = ka_in > kb_in;
    const QG &A = sw ? B_in : A_in;
    const QG &B = sw ? A_in : B_in;
    const int ka = sw ? kb_in : ka_in, kb = sw ? ka_in : kb_in;
    const Vec3<float> d = B.c - A.c;
    if (ka == K_SPHERE) {
        if (kb == K_SPHERE) return classify(sqrtf(dot(d, d)) - A.h0 - B.h0);
        Vec3<float> p = mk<float>(-dot(d, B.x), -dot(d, B.y), -dot(d, B.z)); /* A's centre in B's frame */
        float dist = (kb == K_BOX) ? point_box(p, B.h0, B.h1, B.h2) : point_cyl(p, B.h0, B.h1);
        return classify(dist - A.h0);
    }
    if (ka == K_BOX && kb == K_BOX) {
        /* face axes of both boxes */
        float ta0 = dot(d, A.x), ta1 = dot(d, A.y), ta2 = dot(d, A.z);
        float g = fabsf(ta0) - A.h0 - box_proj(B, A.x);
        g = fmaxf(g, fabsf(ta1) - A.h1 - box_proj(B, A.y));
        g = fmaxf(g, fabsf(ta2) - A.h2 - box_proj(B, A.z));
        float tb0 = dot(d, B.x), tb1 = dot(d, B.y), tb2 = dot(d, B.z);
        g = fmaxf(g, fabsf(tb0) - B.h0 - box_proj(A, B.x));
        g = fmaxf(g, fabsf(tb1) - B.h1 - box_proj(A, B.y));
        g = fmaxf(g, fabsf(tb2) - B.h2 - box_proj(A, B.z));
        if (g > eps) return SEP;
        if (fabsf(ta0) < A.h0 - eps && fabsf(ta1) < A.h1 - eps && fabsf(ta2) < A.h2 - eps) return HIT; /* B's centre in A */
        if (fabsf(tb0) < B.h0 - eps && fabsf(tb1) < B.h1 - eps && fabsf(tb2) < B.h2 - eps) return HIT; /* A's centre in B */
        return NEED;
    }
    if (ka == K_BOX) { /* box A, cylinder B (axis B.z, radius B.h0, half height B.h1) */
        float t0 = dot(d, A.x), t1 = dot(d, A.y), t2 = dot(d, A.z);
        float g = fabsf(t0) - A.h0 - cyl_proj(B.z, B.h0, B.h1, A.x);
        g = fmaxf(g, fabsf(t1) - A.h1 - cyl_proj(B.z, B.h0, B.h1, A.y));
        g = fmaxf(g, fabsf(t2) - A.h2 - cyl_proj(B.z, B.h0, B.h1, A.z));
        float tz = dot(d, B.z);
        g = fmaxf(g, fabsf(tz) - B.h1 - box_proj(A, B.z));
        Vec3<float> dp = d - B.z * tz; /* radial direction from the cylinder axis to the box centre */
        float rho2 = dot(dp, dp);
        float rho = sqrtf(rho2);
        if (rho2 > 1e-12f) {
            Vec3<float> n = dp * (1.0f / rho);
            g = fmaxf(g, rho - B.h0 - box_proj(A, n));
        }
        if (g > eps) return SEP;
        if (fabsf(t0) < A.h0 - eps && fabsf(t1) < A.h1 - eps && fabsf(t2) < A.h2 - eps) return HIT; /* cyl centre in box */
        if (fabsf(tz) < B.h1 - eps && rho < B.h0 - eps) return HIT;                                   /* box centre in cyl */
        {
            /* the cylinder's axis segment passes through the box shrunk by eps (the usual way an arm link
             * meets a table slab: its centre is outside, its axis crosses) - slab clipping in A's frame */
            const float tt[3] = {t0, t1, t2};
            const float uu[3] = {dot(B.z, A.x), dot(B.z, A.y), dot(B.z, A.z)};
            const float gg[3] = {A.h0 - eps, A.h1 - eps, A.h2 - eps};
            float lo = -(B.h1 - eps), hi = B.h1 - eps;
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                if (gg[i] <= 0.f) ok = false;
                if (fabsf(uu[i]) < 1e-6f) {
                    if (fabsf(tt[i]) >= gg[i]) ok = false;
                } else {
                    const float inv = 1.0f / uu[i];
                    const float sa = (-gg[i] - tt[i]) * inv, sb = (gg[i] - tt[i]) * inv;
                    lo = fmaxf(lo, fminf(sa, sb));
                    hi = fminf(hi, fmaxf(sa, sb));
                }
            }
            if (ok && lo < hi) return HIT;
            /* the cylinder's extreme point towards each face plane of the box (support point of the
             * cylinder shrunk by eps) lies inside the shrunk box: a rim dipping into a slab */
            if (ok) {
                const float rr = B.h0 - eps, hh = B.h1 - eps;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const float sg = tt[i] > 0.f ? -1.0f : 1.0f; /* n = sg * e_i */
                    const float nu = sg * uu[i];
                    const float pl2 = 1.0f - uu[i] * uu[i];
                    const float ax = nu >= 0.f ? hh : -hh;
                    float q[3] = {tt[0] + uu[0] * ax, tt[1] + uu[1] * ax, tt[2] + uu[2] * ax};
                    if (pl2 > 1e-6f) {
                        const float k = rr / sqrtf(pl2);
#pragma unroll
                        for (int c = 0; c < 3; c++) q[c] += k * (sg * ((c == i ? 1.0f : 0.0f) - uu[i] * uu[c]));
                    }
                    if (fabsf(q[0]) < gg[0] && fabsf(q[1]) < gg[1] && fabsf(q[2]) < gg[2]) return HIT;
                    /* thin slabs: the extreme point overshoots to the far side.  From a point P of the axis
                     * segment walk radially (perpendicular to the axis, towards the box's mid-plane x_i = 0)
                     * until that plane; if it is reached within the radius the landing point belongs to the
                     * cylinder, and it is inside the box if its other two coordinates are. */
                    if (pl2 > 1e-6f) {
                        const float ipl = 1.0f / pl2;
#pragma unroll
                        for (int c3 = 0; c3 < 3; c3++) {
                            const float sa = c3 == 0 ? 0.f : (c3 == 1 ? hh : -hh);
                            const float P[3] = {tt[0] + uu[0] * sa, tt[1] + uu[1] * sa, tt[2] + uu[2] * sa};
                            if (P[i] * P[i] * ipl < rr * rr) { /* radial distance to the plane < radius */
                                const float f = -P[i] * ipl;    /* q = P + f * (e_i - u_i u) */
                                bool in = true;
#pragma unroll
                                for (int c = 0; c < 3; c++) {
                                    if (c == i) continue;
                                    const float qc = P[c] - f * uu[i] * uu[c];
                                    in = in && fabsf(qc) < gg[c];
                                }
                                if (in) return HIT;
                            }
                        }
                    }
                }
            }
        }
        return NEED;
    }
    /* cylinder - cylinder */
    {
        float ta = dot(d, A.z), tb = dot(d, B.z);
        float g = fabsf(ta) - A.h1 - cyl_proj(B.z, B.h0, B.h1, A.z);
        g = fmaxf(g, fabsf(tb) - B.h1 - cyl_proj(A.z, A.h0, A.h1, B.z));
        Vec3<float> n = cross(A.z, B.z);
        float n2 = dot(n, n);
        if (n2 > 1e-6f) { /* common perpendicular of the two axes: both cylinders project to their radius */
            float inv = 1.0f / sqrtf(n2);
            g = fmaxf(g, fabsf(dot(d, n)) * inv - A.h0 - B.h0);
        }
        Vec3<float> da = d - A.z * ta; /* radial direction of A towards B's centre */
        float ra2 = dot(da, da), ra = sqrtf(ra2);
        if (ra2 > 1e-12f) {
            Vec3<float> m = da * (1.0f / ra);
            g = fmaxf(g, ra - A.h0 - cyl_proj(B.z, B.h0, B.h1, m));
        }
        Vec3<float> db = d - B.z * tb;
        float rb2 = dot(db, db), rb = sqrtf(rb2);
        if (rb2 > 1e-12f) {
            Vec3<float> m = db * (1.0f / rb);
            g = fmaxf(g, rb - B.h0 - cyl_proj(A.z, A.h0, A.h1, m));
        }
        if (g > eps) return SEP;
        if (fabsf(ta) < A.h1 - eps && ra < A.h0 - eps) return HIT; /* B's centre in A */
        if (fabsf(tb) < B.h1 - eps && rb < B.h0 - eps) return HIT; /* A's centre in B */
        return NEED;
    }
}


/* ------------------------------------------------------------------ meshes: AABB tree in mesh-local frame */
struct BvhNode {
    float lo[3];
    int a; /* inner: left child | leaf: first triangle */
    float hi[3];
    int b; /* inner: -right child - 1 ... encoded as negative | leaf: triangle count (> 0) */
};

template <typename T> struct MeshRef {
    const BvhNode *nodes; /* root = 0 */
    const T *tris;        /* 9 values per triangle, mesh-local */
    T brad;               /* bounding radius about the mesh-local origin of `centre` */
    T cx, cy, cz;         /* bounding-sphere centre, mesh-local */
};

/* triangles are stored 12 values apart (9 used) so that a float triangle is three aligned 16-byte loads,
 * nodes are 32 bytes: two 16-byte loads */
#define TMK_TRI_STRIDE 12
template <typename T> TMK_HD void load_tri(const T *tris, int tri, Vec3<T> &a, Vec3<T> &b, Vec3<T> &c) {
    const T *tp = tris + TMK_TRI_STRIDE * (size_t)tri;
    a = mk<T>(tp[0], tp[1], tp[2]);
    b = mk<T>(tp[3], tp[4], tp[5]);
    c = mk<T>(tp[6], tp[7], tp[8]);
}
#ifdef __CUDA_ARCH__
template <> __device__ __forceinline__ void load_tri<float>(const float *tris, int tri, Vec3<float> &a, Vec3<float> &b,
                                                            Vec3<float> &c) {
    const float4 *t4 = reinterpret_cast<const float4 *>(tris) + 3 * (size_t)tri;
    const float4 u = __ldg(t4), v = __ldg(t4 + 1), w = __ldg(t4 + 2);
    a = mk<float>(u.x, u.y, u.z);
    b = mk<float>(u.w, v.x, v.y);
    c = mk<float>(v.z, v.w, w.x);
}
#endif
TMK_HD BvhNode load_node(const BvhNode *nodes, int i) {
#ifdef __CUDA_ARCH__
    const float4 *n4 = reinterpret_cast<const float4 *>(nodes + i);
    const float4 u = __ldg(n4), v = __ldg(n4 + 1);
    BvhNode nd;
    nd.lo[0] = u.x; nd.lo[1] = u.y; nd.lo[2] = u.z; nd.a = __float_as_int(u.w);
    nd.hi[0] = v.x; nd.hi[1] = v.y; nd.hi[2] = v.z; nd.b = __float_as_int(v.w);
    return nd;
#else
    return nodes[i];
#endif
}


/* quick verdict of a triangle (points in S's local frame) against a box / cylinder: SEP by one of
 * the shape's own axes, the triangle's plane normal or (cylinder) the radial direction towards the
 * triangle; HIT when a vertex lies inside the shape by more than the allowance; else NEED. */
template <typename T> TMK_HD int quick_tri(const Shape<T> &S, Vec3<T> p0, Vec3<T> p1, Vec3<T> p2) {
    const T eps = Num<T>::eps();
    const T zlo = t_min(p0.z, t_min(p1.z, p2.z)), zhi = t_max(p0.z, t_max(p1.z, p2.z));
    Vec3<T> n = cross(p1 - p0, p2 - p0);
    const T n2 = dot(n, n);
    if (S.kind == K_BOX) {
        if (zlo > S.c + eps || zhi < -S.c - eps) return SEP;
        if (t_min(p0.x, t_min(p1.x, p2.x)) > S.a + eps || t_max(p0.x, t_max(p1.x, p2.x)) < -S.a - eps) return SEP;
        if (t_min(p0.y, t_min(p1.y, p2.y)) > S.b + eps || t_max(p0.y, t_max(p1.y, p2.y)) < -S.b - eps) return SEP;
        if (n2 > Num<T>::tiny()) {
            const T inv = T(1) / t_sqrt(n2);
            const T d = t_abs(dot(n, p0)) * inv;
            const T proj = (S.a * t_abs(n.x) + S.b * t_abs(n.y) + S.c * t_abs(n.z)) * inv;
            if (d > proj + eps) return SEP;
        }
        const T hx = S.a - eps, hy = S.b - eps, hz = S.c - eps;
        if (t_abs(p0.x) < hx && t_abs(p0.y) < hy && t_abs(p0.z) < hz) return HIT;
        if (t_abs(p1.x) < hx && t_abs(p1.y) < hy && t_abs(p1.z) < hz) return HIT;
        if (t_abs(p2.x) < hx && t_abs(p2.y) < hy && t_abs(p2.z) < hz) return HIT;
        return NEED;
    }
    /* cylinder: S.a = radius, S.b = half height, axis z */
    if (zlo > S.b + eps || zhi < -S.b - eps) return SEP;
    if (n2 > Num<T>::tiny()) {
        const T inv = T(1) / t_sqrt(n2);
        const T d = t_abs(dot(n, p0)) * inv;
        const T proj = (S.b * t_abs(n.z) + S.a * t_sqrt(n.x * n.x + n.y * n.y)) * inv;
        if (d > proj + eps) return SEP;
    }
    {
        const T cx = p0.x + p1.x + p2.x, cy = p0.y + p1.y + p2.y;
        const T c2 = cx * cx + cy * cy;
        if (c2 > Num<T>::tiny()) {
            const T inv = T(1) / t_sqrt(c2);
            const T mx = cx * inv, my = cy * inv;
            const T lo = t_min(p0.x * mx + p0.y * my, t_min(p1.x * mx + p1.y * my, p2.x * mx + p2.y * my));
            if (lo > S.a + eps) return SEP;
        }
    }
    const T r2 = (S.a - eps) * (S.a - eps), hz = S.b - eps;
    if (S.a > eps) {
        if (t_abs(p0.z) < hz && p0.x * p0.x + p0.y * p0.y < r2) return HIT;
        if (t_abs(p1.z) < hz && p1.x * p1.x + p1.y * p1.y < r2) return HIT;
        if (t_abs(p2.z) < hz && p2.x * p2.x + p2.y * p2.y < r2) return HIT;
    }
    return NEED;
}

/* what to do with a triangle the FP32 certificates cannot decide: by default the whole pair becomes
 * UNC; the batch kernels pass a functor that records the single triangle for the exact re-check */
struct NoTri {
    static constexpr bool enabled = false;
    TMK_HD void operator()(int) const {}
};

/* one triangle of a mesh (pose XM) against shape S (pose XS) */
template <typename T>
TMK_HD int test_mesh_tri_shape(const MeshRef<T> &M, const Xf<T> &XM, const Shape<T> &S, const Xf<T> &XS, int tri) {
    Xf<T> m_in_s = xrel(XS, XM);
    Mat3<T> Rms = q2m(m_in_s.r);
    Vec3<T> t0, t1, t2;
    load_tri(M.tris, tri, t0, t1, t2);
    Vec3<T> p0 = mv(Rms, t0) + m_in_s.v;
    Vec3<T> p1 = mv(Rms, t1) + m_in_s.v;
    Vec3<T> p2 = mv(Rms, t2) + m_in_s.v;
    return test_tri_vs_shape(S, p0, p1, p2);
}

/* a convex shape S against a mesh, in the mesh frame: what the tree walk needs per query */
template <typename T> struct MeshQuery {
    Mat3<T> Rms;       /* mesh frame -> S frame */
    Vec3<T> tms;       /* ... and its translation */
    Vec3<T> c;         /* S's centre in the mesh frame */
    T ex, ey, ez;      /* S's extents along the mesh axes (padded) */
    T rad2;            /* (bounding radius of S + pad)^2 */
    TMK_HD void init(const Xf<T> &XM, const Shape<T> &S, const Xf<T> &XS, T s_brad) {
        Xf<T> s_in_m = xrel(XM, XS); /* S in mesh frame */
        Xf<T> m_in_s = xrel(XS, XM); /* mesh in S frame */
        Rms = q2m(m_in_s.r);
        tms = m_in_s.v;
        c = s_in_m.v;
        const T pad = Num<T>::eps() * T(4);
        T rad = s_brad + pad;
        rad2 = rad * rad;
        /* extents of S along the mesh axes (rows of Rms = S's axes in the mesh frame, transposed) */
        if (S.kind == K_SPHERE) {
            ex = ey = ez = S.a + pad;
        } else if (S.kind == K_BOX) {
            ex = S.a * t_abs(Rms.m00) + S.b * t_abs(Rms.m10) + S.c * t_abs(Rms.m20) + pad;
            ey = S.a * t_abs(Rms.m01) + S.b * t_abs(Rms.m11) + S.c * t_abs(Rms.m21) + pad;
            ez = S.a * t_abs(Rms.m02) + S.b * t_abs(Rms.m12) + S.c * t_abs(Rms.m22) + pad;
        } else { /* cylinder axis in the mesh frame = third row of Rms */
            const T ux = Rms.m20, uy = Rms.m21, uz = Rms.m22;
            ex = S.b * t_abs(ux) + S.a * t_sqrt(t_max(T(0), T(1) - ux * ux)) + pad + T(1e-6) * S.a;
            ey = S.b * t_abs(uy) + S.a * t_sqrt(t_max(T(0), T(1) - uy * uy)) + pad + T(1e-6) * S.a;
            ez = S.b * t_abs(uz) + S.a * t_sqrt(t_max(T(0), T(1) - uz * uz)) + pad + T(1e-6) * S.a;
        }
    }
    /* S's axis-aligned bounds in the mesh frame against the node, then its bounding sphere */
    TMK_HD bool overlaps(const BvhNode &nd) const {
        if (c.x - ex > T(nd.hi[0]) || c.x + ex < T(nd.lo[0]) || c.y - ey > T(nd.hi[1]) || c.y + ey < T(nd.lo[1]) ||
            c.z - ez > T(nd.hi[2]) || c.z + ez < T(nd.lo[2]))
            return false;
        T dx = t_max(t_max(T(nd.lo[0]) - c.x, c.x - T(nd.hi[0])), T(0));
        T dy = t_max(t_max(T(nd.lo[1]) - c.y, c.y - T(nd.hi[1])), T(0));
        T dz = t_max(t_max(T(nd.lo[2]) - c.z, c.z - T(nd.hi[2])), T(0));
        return !(dx * dx + dy * dy + dz * dz > rad2);
    }
    /* triangle `tri` of the mesh in S's frame */
    TMK_HD void tri_in_s(const T *tris, int tri, Vec3<T> &p0, Vec3<T> &p1, Vec3<T> &p2) const {
        Vec3<T> t0, t1, t2;
        load_tri(tris, tri, t0, t1, t2);
        p0 = mv(Rms, t0) + tms;
        p1 = mv(Rms, t1) + tms;
        p2 = mv(Rms, t2) + tms;
    }
};

/* shape S (sphere/box/cyl, pose XS) against mesh (pose XM) */
/* UncTri: what to do with a triangle the FP32 tests cannot decide.  NeedTri: what to do with a
 * triangle the quick certificates leave open - by default it is decided here (GJK); the wavefront
 * kernels record it instead so that the traversal stays short and uniform. */
template <typename T, class UncTri = NoTri, class NeedTri = NoTri>
TMK_HD int test_mesh_shape(const MeshRef<T> &M, const Xf<T> &XM, const Shape<T> &S, const Xf<T> &XS, T s_brad,
                           unsigned *n_leaf = nullptr, UncTri on_unc_tri = UncTri(), NeedTri on_need_tri = NeedTri()) {
    MeshQuery<T> Q;
    Q.init(XM, S, XS, s_brad);
    int stack[32];
    int sp = 0;
    stack[sp++] = 0;
    int result = SEP;
    while (sp > 0) {
        BvhNode nd = load_node(M.nodes, stack[--sp]);
        if (!Q.overlaps(nd)) continue;
        if (nd.b > 0) {
            for (int k = 0; k < nd.b; k++) {
                Vec3<T> p0, p1, p2;
                Q.tri_in_s(M.tris, nd.a + k, p0, p1, p2);
                if (n_leaf) (*n_leaf)++;
                int r;
                if (S.kind == K_SPHERE) r = test_tri_vs_shape(S, p0, p1, p2);
                else {
                    r = quick_tri(S, p0, p1, p2);
                    if (r == NEED) {
                        if (NeedTri::enabled) { on_need_tri(nd.a + k); r = SEP; }
                        else r = test_tri_vs_shape(S, p0, p1, p2);
                    }
                }
                if (r == HIT) return HIT;
                if (r == UNC) {
                    if (UncTri::enabled) on_unc_tri(nd.a + k);
                    else result = UNC;
                }
            }
        } else {
            if (sp + 2 <= 32) {
                stack[sp++] = nd.a;
                stack[sp++] = -nd.b - 1;
            } else {
                result = UNC; /* cannot happen for trees built by the host (depth <= 30) */
            }
        }
    }
    return result;
}

/* triangle-triangle, exact path only (double): GJK of two flat sets in world coordinates */
template <typename T>
TMK_HD int test_tri_tri(Vec3<T> a0, Vec3<T> a1, Vec3<T> a2, Vec3<T> b0, Vec3<T> b1, Vec3<T> b2) {
    /* express as "A = degenerate box" is not possible; run GJK with A = triangle by shifting frames:
     * use B as GjkB triangle and A as a triangle through a tiny adaptor: swap roles via support of A */
    Vec3<T> v = a0 - b0;
    if (dot(v, v) < T(1e-24)) v = mk<T>(T(1), T(0), T(0));
    GjkB<T> A, B;
    A.kind = K_TRI; A.p0 = a0; A.p1 = a1; A.p2 = a2;
    B.kind = K_TRI; B.p0 = b0; B.p1 = b1; B.p2 = b2;
    Vec3<T> W[4];
    int n = 0;
    for (int it = 0; it < 64; it++) {
        Vec3<T> w = sup_B(A, -v) - sup_B(B, v);
        T vv = dot(v, v), vw = dot(v, w);
        if (vw > T(0)) return SEP;
        if (n > 0 && vv <= T(1e-30)) return HIT;
        for (int k = 0; k < n; k++) {
            Vec3<T> d = w - W[k];
            if (dot(d, d) <= T(1e-28) * (T(1) + dot(w, w))) return vv > T(1e-24) ? SEP : HIT;
        }
        W[n++] = w;
        Vec3<T> nv;
        bool inside;
        n = simplex_solve(W, n, nv, inside);
        if (inside) return HIT;
        T nn = dot(nv, nv);
        if (it > 0 && nn >= vv) return vv > T(1e-24) ? SEP : HIT;
        v = nv;
    }
    return dot(v, v) > T(1e-24) ? SEP : HIT;
}

/* ------------------------------------------------------------------ OMPL test order
 * k-th interior index (k >= 0) of DiscreteMotionValidator's FIFO bisection over 1..nd-1.
 * The bisection tree is size-balanced, so every level but the last is full: the k-th node in
 * breadth-first order is node j = k+1-2^L of level L = floor(log2(k+1)); descend from the root
 * counting how many level-L nodes fall in the left subtree. */
TMK_HD int ompl_level_count(int m, int d) {
    /* nodes at relative depth d in a balanced bisection tree over m indices */
    long long full = 1ll << d;
    long long c = (long long)m - (full - 1);
    if (c < 0) c = 0;
    if (c > full) c = full;
    return (int)c;
}
TMK_HD int ompl_mid(int nd, int k) {
    int m = nd - 1, base = 1;
#ifdef __CUDA_ARCH__
    int L = 31 - __clz(k + 1);
#else
    int L = 31 - __builtin_clz((unsigned)(k + 1));
#endif
    int j = k + 1 - (1 << L);
    for (int d = L; d > 0; d--) {
        int ml = (m - 1) / 2;
        int cl = ompl_level_count(ml, d - 1);
        if (j < cl) m = ml;
        else { j -= cl; base += ml + 1; m = m - 1 - ml; }
    }
    return base + (m - 1) / 2;
}


} /* namespace tmk */
