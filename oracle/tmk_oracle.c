/* tmk_oracle.c - CPU oracle for the motion-planning validity check.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (tmkit_b200/csrc -> libtmkcl.so) never links or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference's implementation of this path lives in Amino -> FCL /
 * libccd -> OMPL, none of which is under /root/reference, none of which is pinned
 * (/root/reference/.siph:5, configure.ac:34) and none of which is installed here; the
 * reference tree holds no golden vector for the path (no TESTS in Makefile.am).  This
 * file is a plain-C, double-precision restatement of the published semantics:
 *
 *   FK          aa_rx_sg_tf call shape           demo/tmpexec.c:322-325, 407-408
 *               (frames topologically sorted; rel = E, E*Rot(axis,q+off) or
 *               E*Trans(axis*(q+off)); abs = abs[parent] * rel; 7 doubles = quaternion
 *               x,y,z,w + translation, class.robray:61-65, lisp/driver.lisp:5)
 *   collision   aa_rx_cl_check call shape        demo/tmpexec.c:409-421
 *               (one object per collision geometry; pairs on the same frame and pairs
 *               whose FRAMES are in the allowed set are skipped; with a set argument
 *               every colliding frame pair is recorded; the set is n_f x n_f symmetric)
 *   allowance   aa_rx_sg_allow_config            demo/tmpexec.c:118
 *   shapes      box = centred full extents (tm-blocks.py:213-217), sphere centred,
 *               cylinder = +z axis with its base at the origin (see scenes.py header),
 *               mesh = triangle soup (surface only, as FCL's BVHModel)
 *   edges       OMPL DiscreteMotionValidator::checkMotion (external):
 *               test s2, nd = ceil(|s2-s1| / seg_len), then FIFO bisection over 1..nd-1
 *
 * Every convex pair is decided with ONE primitive, gjk_sep(A, B, t): "is the distance
 * between the cores of A and B larger than t?".  From it:
 *     collide(A,B)  = !gjk_sep(A, B, rA + rB)                      (the verdict)
 *     FREE          =  gjk_sep(A, B, rA + rB + tol)                (separation > tol)
 *     HIT           = !gjk_sep(A-, B-, rA- + rB-)   A-,B- = shapes eroded by tol/2
 *     BAND          = neither  (contact band; excluded from bit-exact comparison)
 * where r is the radius of a sphere (its core is a point).
 *
 * Fast mode (what bench.py times as the CPU baseline): the scan stops at the first HIT
 * (early_out), every geometry's world shape is built once per configuration, and bounding spheres - of
 * geometries, of whole meshes, of single triangles in their mesh frame - skip pairs that are more than tol
 * apart.  A skipped pair is FREE by construction, so tri-states and verdicts are those of the plain scan
 * (tests/test_oracle_golden.py compares the two).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SPHERE 0
#define ORC_BOX 1
#define ORC_CYLINDER 2
#define ORC_MESH 3
#define ORC_TRI 4

#define ORC_FREE 0
#define ORC_HIT 1
#define ORC_BAND 2

typedef struct {
    int n_f, n_q;
    const int *parent;  /* [n_f]  parent < child; -1 = root */
    const int *type;    /* 0 fixed, 1 revolute, 2 prismatic */
    const double *E;    /* [n_f*7] */
    const double *axis; /* [n_f*3] */
    const double *offset;
    const int *qidx;
    int n_g;
    const int *g_frame;
    const int *g_shape;
    const double *g_dim; /* [n_g*3] */
    const int *g_mesh;
    int n_mesh;
    const int *mesh_vstart;
    const int *mesh_tstart;
    const double *verts; /* [nv*3] */
    const int *tris;     /* [nt*3], indices local to the mesh */
    const uint8_t *allowed; /* [n_f*n_f] */
    /* optional (may be NULL): bounding sphere of every mesh in its own frame - centre [n_mesh*3], radius [n_mesh].
     * Only used to skip work whose outcome is FREE anyway (the fast mode SURVEY.md 8d asks the CPU baseline to have) */
    const double *mesh_bc;
    const double *mesh_br;
    /* optional: centroid [nt*3] and bounding radius about it [nt] of every triangle in its mesh frame
     * (rows follow `tris`): triangles are culled there before they are transformed */
    const double *tri_c;
    const double *tri_b;
} orc_scene;

/* ------------------------------------------------------------------ vector / quaternion */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void sub3(const double *a, const double *b, double *c) { c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2]; }

static void qmul(const double *a, const double *b, double *c) {
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    c[0] = x; c[1] = y; c[2] = z; c[3] = w;
}
static void qrot(const double *q, const double *v, double *out) {
    double t[3], u[3];
    cross3(q, v, t);
    t[0] *= 2; t[1] *= 2; t[2] *= 2;
    cross3(q, t, u);
    out[0] = v[0] + q[3] * t[0] + u[0];
    out[1] = v[1] + q[3] * t[1] + u[1];
    out[2] = v[2] + q[3] * t[2] + u[2];
}
/* (r1,v1)*(r2,v2) = (r1 r2, rot(r1,v2)+v1)   SURVEY.md A.2 */
static void tfmul(const double *a, const double *b, double *c) {
    double r[4], v[3];
    qmul(a, b, r);
    qrot(a, b + 4, v);
    c[0] = r[0]; c[1] = r[1]; c[2] = r[2]; c[3] = r[3];
    c[4] = v[0] + a[4]; c[5] = v[1] + a[5]; c[6] = v[2] + a[6];
}
static void quat2mat(const double *q, double *R) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

/* ------------------------------------------------------------------ FK */
void orc_fk(const orc_scene *s, const double *q, double *TF_rel, double *TF_abs) {
    for (int i = 0; i < s->n_f; i++) {
        const double *E = s->E + 7 * i;
        double *rel = TF_rel + 7 * i;
        if (s->type[i] == 0) {
            memcpy(rel, E, 7 * sizeof(double));
        } else {
            double a = q[s->qidx[i]] + s->offset[i];
            const double *ax = s->axis + 3 * i;
            if (s->type[i] == 1) {
                double h = 0.5 * a, sn = sin(h), cs = cos(h);
                double n = sqrt(dot3(ax, ax));
                double qa[4] = {ax[0] / n * sn, ax[1] / n * sn, ax[2] / n * sn, cs};
                qmul(E, qa, rel);
                rel[4] = E[4]; rel[5] = E[5]; rel[6] = E[6];
            } else {
                double d[3] = {ax[0] * a, ax[1] * a, ax[2] * a}, r[3];
                qrot(E, d, r);
                rel[0] = E[0]; rel[1] = E[1]; rel[2] = E[2]; rel[3] = E[3];
                rel[4] = E[4] + r[0]; rel[5] = E[5] + r[1]; rel[6] = E[6] + r[2];
            }
        }
        if (s->parent[i] < 0) memcpy(TF_abs + 7 * i, rel, 7 * sizeof(double));
        else tfmul(TF_abs + 7 * s->parent[i], rel, TF_abs + 7 * i);
    }
}

/* ------------------------------------------------------------------ convex shapes */
typedef struct {
    int kind;       /* ORC_SPHERE (core = point), ORC_BOX, ORC_CYLINDER (centred), ORC_TRI */
    double c[3];    /* world centre */
    double R[9];    /* local -> world */
    double h[3];    /* box half extents | cylinder: radius, half height */
    double tri[9];  /* world vertices */
    double rad;     /* sphere radius (added to the point core) */
    double bound;   /* bounding-sphere radius about c (tri: about centroid) */
} cvx;

static void support(const cvx *s, const double *d, double *out) {
    switch (s->kind) {
    case ORC_SPHERE:
        out[0] = s->c[0]; out[1] = s->c[1]; out[2] = s->c[2];
        return;
    case ORC_BOX: {
        const double *R = s->R;
        double l[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
                       R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
        double p[3] = {l[0] >= 0 ? s->h[0] : -s->h[0], l[1] >= 0 ? s->h[1] : -s->h[1], l[2] >= 0 ? s->h[2] : -s->h[2]};
        for (int i = 0; i < 3; i++) out[i] = s->c[i] + R[3 * i] * p[0] + R[3 * i + 1] * p[1] + R[3 * i + 2] * p[2];
        return;
    }
    case ORC_CYLINDER: {
        const double *R = s->R;
        double l[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
                       R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
        double rho = sqrt(l[0] * l[0] + l[1] * l[1]);
        double p[3] = {0, 0, l[2] >= 0 ? s->h[1] : -s->h[1]};
        if (rho > 0) { p[0] = s->h[0] * l[0] / rho; p[1] = s->h[0] * l[1] / rho; }
        for (int i = 0; i < 3; i++) out[i] = s->c[i] + R[3 * i] * p[0] + R[3 * i + 1] * p[1] + R[3 * i + 2] * p[2];
        return;
    }
    default: {
        int b = 0;
        double best = dot3(s->tri, d);
        for (int k = 1; k < 3; k++) {
            double v = dot3(s->tri + 3 * k, d);
            if (v > best) { best = v; b = k; }
        }
        out[0] = s->tri[3 * b]; out[1] = s->tri[3 * b + 1]; out[2] = s->tri[3 * b + 2];
        return;
    }
    }
}

/* ------------------------------------------------------------------ closest point of a simplex to the origin
 * (Voronoi-region walk).  Input W[n][3]; output: v = closest point, W compacted to the
 * supporting sub-simplex, returns its size; returns 4 iff the origin is inside a
 * non-degenerate tetrahedron (v = 0). */
static int closest_seg(double W[4][3], double *v) {
    double ab[3];
    sub3(W[1], W[0], ab);
    double t = -dot3(W[0], ab), den = dot3(ab, ab);
    if (t <= 0 || den <= 0) { memcpy(v, W[0], 24); return 1; }
    if (t >= den) { memcpy(v, W[1], 24); memcpy(W[0], W[1], 24); return 1; }
    t /= den;
    for (int i = 0; i < 3; i++) v[i] = W[0][i] + t * ab[i];
    return 2;
}

/* closest point on triangle (a,b,c) to origin; writes v and mask of used vertices */
static int closest_tri_mask(const double *a, const double *b, const double *c, double *v) {
    double ab[3], ac[3];
    sub3(b, a, ab); sub3(c, a, ac);
    double d1 = -dot3(ab, a), d2 = -dot3(ac, a);
    if (d1 <= 0 && d2 <= 0) { memcpy(v, a, 24); return 1; }
    double d3 = -dot3(ab, b), d4 = -dot3(ac, b);
    if (d3 >= 0 && d4 <= d3) { memcpy(v, b, 24); return 2; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double t = d1 / (d1 - d3);
        for (int i = 0; i < 3; i++) v[i] = a[i] + t * ab[i];
        return 3;
    }
    double d5 = -dot3(ab, c), d6 = -dot3(ac, c);
    if (d6 >= 0 && d5 <= d6) { memcpy(v, c, 24); return 4; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double t = d2 / (d2 - d6);
        for (int i = 0; i < 3; i++) v[i] = a[i] + t * ac[i];
        return 5;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int i = 0; i < 3; i++) v[i] = b[i] + t * (c[i] - b[i]);
        return 6;
    }
    double den = 1.0 / (va + vb + vc);
    double s = vb * den, t = vc * den;
    for (int i = 0; i < 3; i++) v[i] = a[i] + ab[i] * s + ac[i] * t;
    return 7;
}

static int compact(double W[4][3], const int *idx, int mask) {
    double T[4][3];
    int n = 0;
    for (int k = 0; k < 3; k++)
        if (mask & (1 << k)) { memcpy(T[n], W[idx[k]], 24); n++; }
    memcpy(W, T, sizeof(double) * 3 * n);
    return n;
}

static int closest_simplex(double W[4][3], int n, double *v) {
    if (n == 1) { memcpy(v, W[0], 24); return 1; }
    if (n == 2) return closest_seg(W, v);
    if (n == 3) {
        int idx[3] = {0, 1, 2};
        int m = closest_tri_mask(W[0], W[1], W[2], v);
        return compact(W, idx, m);
    }
    /* tetrahedron */
    static const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
    double best = INFINITY, bv[3] = {0, 0, 0};
    int bmask = 0, bf = -1, degenerate = 0, outside_any = 0;
    double scale = 0;
    for (int k = 0; k < 4; k++) scale = fmax(scale, fmax(fabs(W[k][0]), fmax(fabs(W[k][1]), fabs(W[k][2]))));
    for (int f = 0; f < 4; f++) {
        const double *a = W[F[f][0]], *b = W[F[f][1]], *c = W[F[f][2]], *d = W[F[f][3]];
        double ab[3], ac[3], n3[3], ad[3];
        sub3(b, a, ab); sub3(c, a, ac); cross3(ab, ac, n3); sub3(d, a, ad);
        double sp = -dot3(a, n3), sd = dot3(ad, n3);
        if (fabs(sd) <= 1e-14 * scale * scale * scale) degenerate = 1;
        if (degenerate || sp * sd < 0) {
            double cv[3];
            int m = closest_tri_mask(a, b, c, cv);
            double dd = dot3(cv, cv);
            outside_any = 1;
            if (dd < best) { best = dd; memcpy(bv, cv, 24); bmask = m; bf = f; }
        }
    }
    if (degenerate) {
        /* flat tetrahedron: never claim containment; use the closest face (all four were evaluated
         * from the point degeneracy was seen, re-evaluate the earlier ones) */
        for (int f = 0; f < 4; f++) {
            double cv[3];
            int m = closest_tri_mask(W[F[f][0]], W[F[f][1]], W[F[f][2]], cv);
            double dd = dot3(cv, cv);
            if (dd < best) { best = dd; memcpy(bv, cv, 24); bmask = m; bf = f; }
        }
    }
    if (!outside_any && !degenerate) { v[0] = v[1] = v[2] = 0; return 4; }
    memcpy(v, bv, 24);
    return compact(W, F[bf], bmask);
}

/* ------------------------------------------------------------------ gjk_sep: dist(core A, core B) > t ? */
static long g_gjk_iters = 0; /* diagnostics (not thread safe; only read in single-thread tests) */

static int gjk_sep(const cvx *A, const cvx *B, double t) {
    double v[3] = {A->c[0] - B->c[0], A->c[1] - B->c[1], A->c[2] - B->c[2]};
    if (A->kind == ORC_TRI) for (int i = 0; i < 3; i++) v[i] = A->tri[i] - B->c[i];
    if (B->kind == ORC_TRI) for (int i = 0; i < 3; i++) v[i] = (A->kind == ORC_TRI ? A->tri[i] : A->c[i]) - B->tri[i];
    if (dot3(v, v) < 1e-24) { v[0] = 1; v[1] = 0; v[2] = 0; }
    double W[4][3];
    int n = 0;
    double tt = t * t;
    for (int it = 0; it < 256; it++) {
        g_gjk_iters++;
        double nv[3] = {-v[0], -v[1], -v[2]}, a[3], b[3], w[3];
        support(A, nv, a);
        support(B, v, b);
        sub3(a, b, w);
        double vv = dot3(v, v), vw = dot3(v, w);
        if (vw > 0 && vw * vw > tt * vv) return 1;            /* lower bound on distance beats t */
        if (n > 0 && vv <= tt) return 0;                      /* upper bound (|v|, v in A-B) is within t */
        if (n > 0 && vv - vw <= 1e-13 * vv) return vv > tt;   /* converged */
        for (int k = 0; k < n; k++) {
            double d[3];
            sub3(w, W[k], d);
            if (dot3(d, d) <= 1e-28 * (1 + dot3(w, w))) return vv > tt; /* no new vertex: converged */
        }
        memcpy(W[n], w, 24);
        n++;
        double nvv[3];
        n = closest_simplex(W, n, nvv);
        if (n == 4) return 0; /* origin inside the simplex */
        double nn = dot3(nvv, nvv);
        if (it > 0 && nn >= vv) return vv > tt; /* no numerical progress */
        memcpy(v, nvv, 24);
        if (nn <= 1e-30) return 0; /* touching */
    }
    return dot3(v, v) > tt;
}

long orc_gjk_iters(void) { return g_gjk_iters; }

/* ------------------------------------------------------------------ shape construction */
static void make_cvx(int shape, const double *dim, const double *tf, cvx *o, double erode) {
    memset(o, 0, sizeof(*o));
    quat2mat(tf, o->R);
    o->c[0] = tf[4]; o->c[1] = tf[5]; o->c[2] = tf[6];
    if (shape == ORC_SPHERE) {
        o->kind = ORC_SPHERE;
        o->rad = dim[0] - erode;
        o->bound = dim[0];
    } else if (shape == ORC_BOX) {
        o->kind = ORC_BOX;
        for (int i = 0; i < 3; i++) o->h[i] = 0.5 * dim[i] - erode;
        o->bound = 0.5 * sqrt(dim[0] * dim[0] + dim[1] * dim[1] + dim[2] * dim[2]);
    } else { /* cylinder: Amino base-at-origin -> centred core shifted +h/2 along local z */
        o->kind = ORC_CYLINDER;
        double hh = 0.5 * dim[0];
        o->h[0] = dim[1] - erode;
        o->h[1] = hh - erode;
        for (int i = 0; i < 3; i++) o->c[i] += o->R[3 * i + 2] * hh;
        o->bound = sqrt(hh * hh + dim[1] * dim[1]);
    }
}

static void make_tri(const double *tf, const double *R, const double *v0, const double *v1, const double *v2, cvx *o) {
    memset(o, 0, sizeof(*o));
    o->kind = ORC_TRI;
    const double *vs[3] = {v0, v1, v2};
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++)
            o->tri[3 * k + i] = tf[4 + i] + R[3 * i] * vs[k][0] + R[3 * i + 1] * vs[k][1] + R[3 * i + 2] * vs[k][2];
    for (int i = 0; i < 3; i++) o->c[i] = (o->tri[i] + o->tri[3 + i] + o->tri[6 + i]) / 3;
    double b = 0;
    for (int k = 0; k < 3; k++) {
        double d[3];
        sub3(o->tri + 3 * k, o->c, d);
        b = fmax(b, dot3(d, d));
    }
    o->bound = sqrt(b);
}

/* tri-state of one convex pair; *exact = the boolean verdict */
static int convex_pair_pre(int sa, const double *da, const double *ta, const cvx *triA, const cvx *preA,
                           int sb, const double *db, const double *tb, const cvx *triB, const cvx *preB, double tol,
                           int *exact) {
    cvx A, B;
    if (triA) A = *triA; else if (preA) A = *preA; else make_cvx(sa, da, ta, &A, 0);
    if (triB) B = *triB; else if (preB) B = *preB; else make_cvx(sb, db, tb, &B, 0);
    double d[3];
    sub3(A.c, B.c, d);
    double reach = A.bound + B.bound + tol;
    if (dot3(d, d) > reach * reach) { *exact = 0; return ORC_FREE; } /* conservative sphere cull */
    if (gjk_sep(&A, &B, A.rad + B.rad + tol)) { *exact = 0; return ORC_FREE; }
    /* erode: solids by tol/2 each; a triangle cannot shrink, so its partner takes the whole tol */
    cvx Ae = A, Be = B;
    double ea = 0.5 * tol, eb = 0.5 * tol;
    if (triA && !triB) { ea = 0; eb = tol; }
    if (triB && !triA) { eb = 0; ea = tol; }
    if (triA && triB) { ea = eb = 0; }
    if (!triA) make_cvx(sa, da, ta, &Ae, ea);
    if (!triB) make_cvx(sb, db, tb, &Be, eb);
    if (!gjk_sep(&Ae, &Be, Ae.rad + Be.rad)) { *exact = 1; return ORC_HIT; }
    *exact = !gjk_sep(&A, &B, A.rad + B.rad);
    return ORC_BAND;
}
static int convex_pair(int sa, const double *da, const double *ta, const cvx *triA,
                       int sb, const double *db, const double *tb, const cvx *triB, double tol, int *exact) {
    return convex_pair_pre(sa, da, ta, triA, NULL, sb, db, tb, triB, NULL, tol, exact);
}

/* world centre of a mesh's bounding sphere */
static void mesh_centre(const orc_scene *s, int m, const double *tf, const double *R, double *c) {
    const double *b = s->mesh_bc + 3 * m;
    for (int i = 0; i < 3; i++) c[i] = tf[4 + i] + R[3 * i] * b[0] + R[3 * i + 1] * b[1] + R[3 * i + 2] * b[2];
}
/* spheres (ca, ra), (cb, rb) farther apart than tol: whatever they contain is FREE */
static int spheres_apart(const double *ca, double ra, const double *cb, double rb, double tol) {
    double d[3];
    sub3(ca, cb, d);
    double reach = ra + rb + tol;
    return dot3(d, d) > reach * reach;
}

/* geometry pair (either side may be a mesh = union of triangles) */
static int geom_pair(const orc_scene *s, int gi, int gj, const double *TF_abs, double tol, int *exact, const cvx *pre) {
    int si = s->g_shape[gi], sj = s->g_shape[gj];
    const double *ti = TF_abs + 7 * s->g_frame[gi], *tj = TF_abs + 7 * s->g_frame[gj];
    const double *di = s->g_dim + 3 * gi, *dj = s->g_dim + 3 * gj;
    if (si != ORC_MESH && sj != ORC_MESH)
        return convex_pair_pre(si, di, ti, NULL, pre ? pre + gi : NULL, sj, dj, tj, NULL, pre ? pre + gj : NULL, tol, exact);
    if (sj != ORC_MESH) { /* put the mesh second */
        int t = gi; gi = gj; gj = t;
        t = si; si = sj; sj = t;
        const double *p = ti; ti = tj; tj = p;
        p = di; di = dj; dj = p;
    }
    int state = ORC_FREE, ex = 0;
    *exact = 0;
    double Rj[9];
    quat2mat(tj, Rj);
    int mj = s->g_mesh[gj];
    const double *vj = s->verts + 3 * s->mesh_vstart[mj];
    const int *fj = s->tris + 3 * s->mesh_tstart[mj];
    int ntj = s->mesh_tstart[mj + 1] - s->mesh_tstart[mj];
    double cj[3] = {0, 0, 0};
    if (s->mesh_bc) mesh_centre(s, mj, tj, Rj, cj);
    if (si != ORC_MESH) {
        cvx A0;
        if (pre) A0 = pre[gi]; else make_cvx(si, di, ti, &A0, 0);
        if (s->mesh_bc && spheres_apart(A0.c, A0.bound, cj, s->mesh_br[mj], tol)) return ORC_FREE;
        /* shape centre in the mesh frame: the cull below is the one convex_pair starts with, moved in front of
         * the vertex transforms (slack 1e-12 for the change of frame; a triangle that slips through is culled or
         * decided by convex_pair as before) */
        double cl[3] = {0, 0, 0};
        const double *tcj = NULL, *tbj = NULL;
        if (s->tri_c) {
            double d0[3] = {A0.c[0] - tj[4], A0.c[1] - tj[5], A0.c[2] - tj[6]};
            for (int i = 0; i < 3; i++) cl[i] = Rj[i] * d0[0] + Rj[3 + i] * d0[1] + Rj[6 + i] * d0[2];
            tcj = s->tri_c + 3 * s->mesh_tstart[mj];
            tbj = s->tri_b + s->mesh_tstart[mj];
        }
        for (int t = 0; t < ntj; t++) {
            if (tcj && spheres_apart(cl, A0.bound, tcj + 3 * t, tbj[t], tol + 1e-12)) continue;
            cvx T;
            make_tri(tj, Rj, vj + 3 * fj[3 * t], vj + 3 * fj[3 * t + 1], vj + 3 * fj[3 * t + 2], &T);
            if (spheres_apart(A0.c, A0.bound, T.c, T.bound, tol)) continue; /* the cull convex_pair starts with */
            int st = convex_pair_pre(si, di, ti, NULL, &A0, ORC_TRI, NULL, NULL, &T, NULL, tol, &ex);
            *exact |= ex;
            if (st == ORC_HIT) return ORC_HIT;
            if (st == ORC_BAND) state = ORC_BAND;
        }
        return state;
    }
    /* mesh vs mesh */
    double Ri[9];
    quat2mat(ti, Ri);
    int mi = s->g_mesh[gi];
    const double *vi = s->verts + 3 * s->mesh_vstart[mi];
    const int *fi = s->tris + 3 * s->mesh_tstart[mi];
    int nti = s->mesh_tstart[mi + 1] - s->mesh_tstart[mi];
    if (s->mesh_bc) {
        double ci[3];
        mesh_centre(s, mi, ti, Ri, ci);
        if (spheres_apart(ci, s->mesh_br[mi], cj, s->mesh_br[mj], tol)) return ORC_FREE;
    }
    for (int a = 0; a < nti; a++) {
        cvx TA;
        make_tri(ti, Ri, vi + 3 * fi[3 * a], vi + 3 * fi[3 * a + 1], vi + 3 * fi[3 * a + 2], &TA);
        if (s->mesh_bc && spheres_apart(TA.c, TA.bound, cj, s->mesh_br[mj], tol)) continue;
        for (int b = 0; b < ntj; b++) {
            cvx TB;
            make_tri(tj, Rj, vj + 3 * fj[3 * b], vj + 3 * fj[3 * b + 1], vj + 3 * fj[3 * b + 2], &TB);
            if (spheres_apart(TA.c, TA.bound, TB.c, TB.bound, tol)) continue;
            int st = convex_pair(ORC_TRI, NULL, NULL, &TA, ORC_TRI, NULL, NULL, &TB, tol, &ex);
            *exact |= ex;
            if (st == ORC_HIT) return ORC_HIT;
            if (st == ORC_BAND) state = ORC_BAND;
        }
    }
    return state;
}

/* one pair of primitives given directly (unit-test entry point).  dims / tf as in the scene;
 * shape ORC_TRI takes 9 doubles of world vertices in dim and ignores tf. */
int orc_pair(int sa, const double *da, const double *ta, int sb, const double *db, const double *tb, double tol,
             int *exact) {
    cvx TA, TB;
    const cvx *pa = NULL, *pb = NULL;
    static const double I7[7] = {0, 0, 0, 1, 0, 0, 0};
    static const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (sa == ORC_TRI) { make_tri(I7, I9, da, da + 3, da + 6, &TA); pa = &TA; }
    if (sb == ORC_TRI) { make_tri(I7, I9, db, db + 3, db + 6, &TB); pb = &TB; }
    return convex_pair(sa, da, ta, pa, sb, db, tb, pb, tol, exact);
}

/* ------------------------------------------------------------------ aa_rx_cl_check restatement
 * TF_abs: n_f x 7.  `skip` (optional, n_g bytes): geometry flagged 1 is "static"; a pair of two
 * static geometries is skipped (batch mode hoists those).  pairset (optional, n_f*n_f bytes)
 * receives max(state) per frame pair (symmetric).  Returns the config's tri-state; *exact = verdict.
 * With early_out the scan stops at the first HIT (aa_rx_cl_check with a NULL set). */
int orc_check(const orc_scene *s, const double *TF_abs, double tol, const uint8_t *skip, uint8_t *pairset,
              int early_out, int *exact) {
    int state = ORC_FREE;
    *exact = 0;
    /* every geometry's world shape once per configuration (not once per pair), and its bounding sphere:
     * pairs whose spheres are more than tol apart are FREE without looking further */
    cvx *pre = (cvx *)malloc(sizeof(cvx) * (size_t)(s->n_g > 0 ? s->n_g : 1));
    for (int i = 0; i < s->n_g; i++) {
        const double *tf = TF_abs + 7 * s->g_frame[i];
        if (s->g_shape[i] != ORC_MESH) make_cvx(s->g_shape[i], s->g_dim + 3 * i, tf, &pre[i], 0);
        else {
            memset(&pre[i], 0, sizeof(cvx));
            pre[i].bound = -1; /* unknown */
            if (s->mesh_bc) {
                double R[9];
                quat2mat(tf, R);
                mesh_centre(s, s->g_mesh[i], tf, R, pre[i].c);
                pre[i].bound = s->mesh_br[s->g_mesh[i]];
            }
        }
    }
    for (int i = 0; i < s->n_g; i++) {
        int fi = s->g_frame[i];
        for (int j = i + 1; j < s->n_g; j++) {
            int fj = s->g_frame[j];
            if (fi == fj || s->allowed[fi * s->n_f + fj]) continue;
            if (skip && skip[i] && skip[j]) continue;
            if (pre[i].bound >= 0 && pre[j].bound >= 0 && spheres_apart(pre[i].c, pre[i].bound, pre[j].c, pre[j].bound, tol))
                continue;
            int ex = 0;
            int st = geom_pair(s, i, j, TF_abs, tol, &ex, pre);
            *exact |= ex;
            if (st != ORC_FREE && pairset) {
                uint8_t *p = pairset + fi * s->n_f + fj, *q = pairset + fj * s->n_f + fi;
                if (*p != ORC_HIT) *p = *q = (uint8_t)st;
            }
            if (st == ORC_HIT) {
                state = ORC_HIT;
                if (early_out) { free(pre); return state; }
            } else if (st == ORC_BAND && state == ORC_FREE) state = ORC_BAND;
        }
    }
    free(pre);
    return state;
}

/* frames that move when the configuration variables sub[] change */
static void moving_geoms(const orc_scene *s, int n_sub, const int *sub, uint8_t *g_static) {
    uint8_t *mov = calloc(s->n_f, 1);
    for (int i = 0; i < s->n_f; i++) {
        int m = s->parent[i] >= 0 ? mov[s->parent[i]] : 0;
        if (s->type[i] != 0)
            for (int k = 0; k < n_sub; k++)
                if (sub[k] == s->qidx[i]) m = 1;
        mov[i] = (uint8_t)m;
    }
    for (int g = 0; g < s->n_g; g++) g_static[g] = !mov[s->g_frame[g]];
    free(mov);
}

typedef struct {
    const orc_scene *s;
    long n, lo, hi;
    int n_sub;
    const int *sub;
    const double *q_base;
    const double *Q;
    long ldQ;
    const double *Q1; /* edges */
    double seg_len, tol;
    int early_out;
    const uint8_t *g_static;
    int static_state, static_exact;
    uint8_t *state, *exact;
    int32_t *first_bad;
    int64_t *n_checked;
} job;

static void eval_config(const job *J, const double *qsub, double *q, double *rel, double *abs_, int *state, int *exact) {
    for (int k = 0; k < J->n_sub; k++) q[J->sub[k]] = qsub[k];
    orc_fk(J->s, q, rel, abs_);
    int ex = 0;
    int st = orc_check(J->s, abs_, J->tol, J->g_static, NULL, J->early_out, &ex);
    /* fold in the hoisted static-static result */
    if (J->static_state == ORC_HIT) st = ORC_HIT;
    else if (J->static_state == ORC_BAND && st == ORC_FREE) st = ORC_BAND;
    *state = st;
    *exact = ex | J->static_exact;
}

static void *config_worker(void *arg) {
    job *J = (job *)arg;
    const orc_scene *s = J->s;
    double *q = malloc(sizeof(double) * (s->n_q + 1)), *rel = malloc(sizeof(double) * 7 * s->n_f),
           *abs_ = malloc(sizeof(double) * 7 * s->n_f);
    memcpy(q, J->q_base, sizeof(double) * s->n_q);
    for (long c = J->lo; c < J->hi; c++) {
        int st, ex;
        eval_config(J, J->Q + c * J->ldQ, q, rel, abs_, &st, &ex);
        J->state[c] = (uint8_t)st;
        J->exact[c] = (uint8_t)ex;
    }
    free(q); free(rel); free(abs_);
    return NULL;
}

/* OMPL DiscreteMotionValidator order over interior indices 1..nd-1 (FIFO bisection). */
static int ompl_order(int nd, int *order) {
    if (nd < 2) return 0;
    int cap = 2 * nd + 4, head = 0, tail = 0, n = 0;
    int *qa = malloc(sizeof(int) * cap), *qb = malloc(sizeof(int) * cap);
    qa[tail] = 1; qb[tail] = nd - 1; tail++;
    while (head < tail) {
        int a = qa[head], b = qb[head];
        head++;
        int mid = (a + b) / 2;
        order[n++] = mid;
        if (a < mid) { qa[tail] = a; qb[tail] = mid - 1; tail++; }
        if (mid < b) { qa[tail] = mid + 1; qb[tail] = b; tail++; }
    }
    free(qa); free(qb);
    return n;
}
int orc_ompl_order(int nd, int *order) { return ompl_order(nd, order); }

int orc_segment_count(int n_sub, const double *a, const double *b, double seg_len) {
    double d = 0;
    for (int k = 0; k < n_sub; k++) d += (b[k] - a[k]) * (b[k] - a[k]);
    int nd = (int)ceil(sqrt(d) / seg_len);
    return nd < 1 ? 1 : nd;
}

static void *edge_worker(void *arg) {
    job *J = (job *)arg;
    const orc_scene *s = J->s;
    double *q = malloc(sizeof(double) * (s->n_q + 1)), *rel = malloc(sizeof(double) * 7 * s->n_f),
           *abs_ = malloc(sizeof(double) * 7 * s->n_f), *qs = malloc(sizeof(double) * J->n_sub);
    memcpy(q, J->q_base, sizeof(double) * s->n_q);
    int *order = NULL, ocap = 0;
    for (long e = J->lo; e < J->hi; e++) {
        const double *a = J->Q + e * J->ldQ, *b = J->Q1 + e * J->ldQ;
        int nd = orc_segment_count(J->n_sub, a, b, J->seg_len);
        if (nd + 2 > ocap) { ocap = 2 * nd + 16; order = realloc(order, sizeof(int) * ocap); }
        int ni = ompl_order(nd, order);
        int state = ORC_FREE, exact = 0, first = -1, band_before = 0;
        int64_t checked = 0;
        for (int k = -1; k < ni; k++) { /* k = -1: the end state s2 */
            if (k < 0) memcpy(qs, b, sizeof(double) * J->n_sub);
            else {
                double t = (double)order[k] / (double)nd;
                for (int i = 0; i < J->n_sub; i++) qs[i] = a[i] + (b[i] - a[i]) * t;
            }
            int st, ex;
            eval_config(J, qs, q, rel, abs_, &st, &ex);
            checked++;
            if (ex && !exact) { exact = 1; }
            if (st == ORC_HIT) {
                if (state != ORC_HIT) { first = band_before ? -2 : k + 1; }
                state = ORC_HIT;
                if (J->early_out) break;
            } else if (st == ORC_BAND) {
                if (state == ORC_FREE) state = ORC_BAND;
                if (state != ORC_HIT) band_before = 1;
            }
        }
        J->state[e] = (uint8_t)state;
        J->exact[e] = (uint8_t)exact;
        if (J->first_bad) J->first_bad[e] = first; /* index in test order (0 = s2); -2 = ambiguous; -1 = none */
        if (J->n_checked) J->n_checked[e] = checked;
    }
    free(q); free(rel); free(abs_); free(qs); free(order);
    return NULL;
}

static void run_jobs(job *proto, long n, int n_threads, void *(*fn)(void *)) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    job jobs[256];
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = *proto;
        jobs[t].lo = n * t / n_threads;
        jobs[t].hi = n * (t + 1) / n_threads;
    }
    if (n_threads == 1) { fn(&jobs[0]); return; }
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, fn, &jobs[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
}

static void prep_static(const orc_scene *s, job *J, uint8_t *g_static, int hoist) {
    J->g_static = NULL;
    J->static_state = ORC_FREE;
    J->static_exact = 0;
    if (!hoist) return;
    moving_geoms(s, J->n_sub, J->sub, g_static);
    /* evaluate static-static pairs once, at q_base */
    double *rel = malloc(sizeof(double) * 7 * s->n_f), *abs_ = malloc(sizeof(double) * 7 * s->n_f);
    orc_fk(s, J->q_base, rel, abs_);
    int st = ORC_FREE;
    for (int i = 0; i < s->n_g; i++)
        for (int j = i + 1; j < s->n_g; j++) {
            int fi = s->g_frame[i], fj = s->g_frame[j];
            if (!(g_static[i] && g_static[j]) || fi == fj || s->allowed[fi * s->n_f + fj]) continue;
            int ex = 0;
            int p = geom_pair(s, i, j, abs_, J->tol, &ex, NULL);
            J->static_exact |= ex;
            if (p == ORC_HIT) st = ORC_HIT;
            else if (p == ORC_BAND && st == ORC_FREE) st = ORC_BAND;
        }
    J->static_state = st;
    J->g_static = g_static;
    free(rel); free(abs_);
}

/* Batched validity.  Q: n x ldQ doubles (row per configuration, first n_sub entries used);
 * the other joints stay at q_base.  state[n] tri-state, exact[n] boolean verdicts.
 * hoist: evaluate static-static pairs once (same result, less work). */
void orc_check_batch(const orc_scene *s, long n, int n_sub, const int *sub, const double *q_base, const double *Q,
                     long ldQ, double tol, int early_out, int hoist, int n_threads, uint8_t *state, uint8_t *exact) {
    job J;
    memset(&J, 0, sizeof(J));
    J.s = s; J.n = n; J.n_sub = n_sub; J.sub = sub; J.q_base = q_base; J.Q = Q; J.ldQ = ldQ;
    J.tol = tol; J.early_out = early_out; J.state = state; J.exact = exact;
    uint8_t *g_static = malloc(s->n_g + 1);
    prep_static(s, &J, g_static, hoist);
    run_jobs(&J, n, n_threads, config_worker);
    free(g_static);
}

void orc_check_edges(const orc_scene *s, long n, int n_sub, const int *sub, const double *q_base, const double *Q0,
                     const double *Q1, long ldQ, double seg_len, double tol, int early_out, int hoist, int n_threads,
                     uint8_t *state, uint8_t *exact, int32_t *first_bad, int64_t *n_checked) {
    job J;
    memset(&J, 0, sizeof(J));
    J.s = s; J.n = n; J.n_sub = n_sub; J.sub = sub; J.q_base = q_base; J.Q = Q0; J.Q1 = Q1; J.ldQ = ldQ;
    J.seg_len = seg_len; J.tol = tol; J.early_out = early_out; J.state = state; J.exact = exact;
    J.first_bad = first_bad; J.n_checked = n_checked;
    uint8_t *g_static = malloc(s->n_g + 1);
    prep_static(s, &J, g_static, hoist);
    run_jobs(&J, n, n_threads, edge_worker);
    free(g_static);
}

/* ------------------------------------------------------------------ independent cross-checks used by tests */

/* 15-axis SAT for two boxes; returns the largest normalised axis gap (negative = overlap on
 * every axis = intersecting).  Near-null cross axes are skipped. */
double orc_sat_box_box(const double *da, const double *ta, const double *db, const double *tb) {
    double Ra[9], Rb[9], t[3], best = -INFINITY;
    quat2mat(ta, Ra); quat2mat(tb, Rb);
    for (int i = 0; i < 3; i++) t[i] = tb[4 + i] - ta[4 + i];
    double ha[3] = {da[0] / 2, da[1] / 2, da[2] / 2}, hb[3] = {db[0] / 2, db[1] / 2, db[2] / 2};
    double axes[15][3];
    int n = 0;
    for (int i = 0; i < 3; i++) { axes[n][0] = Ra[i]; axes[n][1] = Ra[3 + i]; axes[n][2] = Ra[6 + i]; n++; }
    for (int i = 0; i < 3; i++) { axes[n][0] = Rb[i]; axes[n][1] = Rb[3 + i]; axes[n][2] = Rb[6 + i]; n++; }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { cross3(axes[i], axes[3 + j], axes[n]); n++; }
    for (int k = 0; k < 15; k++) {
        double L2 = dot3(axes[k], axes[k]);
        if (L2 < 1e-16) continue;
        double L = sqrt(L2), ra = 0, rb = 0;
        for (int i = 0; i < 3; i++) {
            ra += ha[i] * fabs(dot3(axes[i], axes[k]));
            rb += hb[i] * fabs(dot3(axes[3 + i], axes[k]));
        }
        double g = (fabs(dot3(t, axes[k])) - ra - rb) / L;
        if (g > best) best = g;
    }
    return best;
}

/* exact distance from a point to a solid primitive (0 inside) - closed forms */
double orc_point_dist(int shape, const double *dim, const double *tf, const double *p) {
    double R[9], d[3], l[3];
    quat2mat(tf, R);
    for (int i = 0; i < 3; i++) d[i] = p[i] - tf[4 + i];
    for (int i = 0; i < 3; i++) l[i] = R[i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2];
    if (shape == ORC_SPHERE) return fmax(0.0, sqrt(dot3(l, l)) - dim[0]);
    if (shape == ORC_BOX) {
        double s = 0;
        for (int i = 0; i < 3; i++) {
            double e = fabs(l[i]) - dim[i] / 2;
            if (e > 0) s += e * e;
        }
        return sqrt(s);
    }
    double hh = dim[0] / 2, z = l[2] - hh;
    double er = sqrt(l[0] * l[0] + l[1] * l[1]) - dim[1], ez = fabs(z) - hh;
    if (er < 0) er = 0;
    if (ez < 0) ez = 0;
    return sqrt(er * er + ez * ez);
}
