/* Workspace-goal inverse kinematics for the motion query (the reference reaches it
 * through robray:motion-plan's :workspace-goal, lisp/python/tmsmtpy.lisp:55-60).
 *
 * Damped least squares over the kinematic chain, every seed iterated to convergence inside ONE launch:
 * one thread per seed, double precision throughout, so the host pays one round trip per goal instead of
 * one per iteration.  The seeds of a goal fit one warp; the warp leaves the loop together (ballot) once
 * enough seeds have converged and the stragglers had their chance. */
#pragma once
#include "dev_math.cuh"

#define TMK_IK_MAXCOL 24 /* = TMK_MP_MAXSUB */

struct IkFrame {
    double E[7];
    double ax[3];
    double off;
    double qfix; /* value of the joint when it is not one of the varied variables */
    int type;    /* 0 fixed, 1 revolute, 2 prismatic */
    int col;     /* column of the varied variable or -1 */
};

struct IkParams {
    double base[7]; /* pose of the chain root */
    double goal[7];
    int n_fr, n_sub, n_seed, max_iter;
    int enough;     /* stop once this many seeds converged ... */
    int min_iter;   /* ... and at least this many iterations ran */
    double tol;     /* convergence: position and rotation error below tol */
};

namespace tmk {

/* solve (A + lambda2 I) x = b, A symmetric positive semi-definite 6x6 (Cholesky, in place) */
__device__ inline bool ik_solve6(double (&A)[6][6], const double *b, double lambda2, double *x) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = A[i][j] + (i == j ? lambda2 : 0.0);
#pragma unroll
            for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
            if (i == j) {
                if (!(s > 0)) return false;
                A[i][i] = sqrt(s);
            } else A[i][j] = s / A[j][j];
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= A[i][k] * y[k];
        y[i] = s / A[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= A[k][i] * x[k];
        x[i] = s / A[i][i];
    }
    return true;
}

/* Q: n_seed x n_sub (in: seeds, out: final iterate); err: n_seed x 2 (position, rotation error of the final
 * iterate); iters: iterations run */
__global__ void __launch_bounds__(32) k_ik_dls(const IkFrame *__restrict__ fr, IkParams P, const double *__restrict__ lo,
                                                const double *__restrict__ hi, double *__restrict__ Q,
                                                double *__restrict__ err, int *__restrict__ iters) {
    const int s = threadIdx.x;
    const bool active = s < P.n_seed;
    double q[TMK_IK_MAXCOL];
    for (int k = 0; k < P.n_sub; k++) q[k] = active ? Q[(size_t)s * P.n_sub + k] : 0.0;
    const Xf<double> base = xload(P.base);
    const Xf<double> goal = xload(P.goal);
    bool done = !active;
    double ep = 1e300, eo = 1e300;
    int it = 0;
    for (; it < P.max_iter; it++) {
        const unsigned n_done = __popc(__ballot_sync(0xffffffffu, done && active));
        if (__all_sync(0xffffffffu, done)) break;
        if ((int)n_done >= P.enough && it >= P.min_iter) break;
        if (done) continue;
        /* FK along the chain; world axis and origin of every varied joint */
        double aw[TMK_IK_MAXCOL][3], org[TMK_IK_MAXCOL][3];
        int kind[TMK_IK_MAXCOL];
        for (int k = 0; k < P.n_sub; k++) kind[k] = 0;
        Xf<double> X = base;
        for (int f = 0; f < P.n_fr; f++) {
            const IkFrame &F = fr[f];
            const Xf<double> E = xload(F.E);
            if (F.type == 0) { X = xmul(X, E); continue; }
            Vec3<double> ax;
            ax.x = F.ax[0]; ax.y = F.ax[1]; ax.z = F.ax[2];
            const double qv = F.col >= 0 ? q[F.col] : F.qfix;
            X = fk_joint<double>(E, ax, F.off, F.type, qv, &X);
            if (F.col >= 0) {
                const double inv = 1.0 / sqrt(ax.x * ax.x + ax.y * ax.y + ax.z * ax.z);
                const Vec3<double> w = qrot(X.r, ax * inv);
                /* a variable that drives several joints of the chain: the columns add up below, so keep
                 * the contributions per joint (the common case is one joint per variable) */
                if (kind[F.col] == 0) {
                    aw[F.col][0] = w.x; aw[F.col][1] = w.y; aw[F.col][2] = w.z;
                    org[F.col][0] = X.v.x; org[F.col][1] = X.v.y; org[F.col][2] = X.v.z;
                    kind[F.col] = F.type;
                } else kind[F.col] = -1; /* handled by the slow path */
            }
        }
        /* pose error: translation, rotation vector of goal * conj(tip) */
        double e[6];
        e[0] = goal.v.x - X.v.x; e[1] = goal.v.y - X.v.y; e[2] = goal.v.z - X.v.z;
        {
            const Quat<double> d = qmul(goal.r, qconj(X.r));
            const double sgn = d.w < 0 ? -1.0 : 1.0;
            const double vn = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
            const double ang = 2.0 * atan2(vn, sgn * d.w);
            const double sc = vn > 1e-12 ? sgn * ang / vn : 2.0 * sgn;
            e[3] = sc * d.x; e[4] = sc * d.y; e[5] = sc * d.z;
        }
        ep = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        eo = sqrt(e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
        if (ep < P.tol && eo < P.tol) { done = true; continue; }
        /* Jacobian columns */
        double J[6][TMK_IK_MAXCOL];
        for (int k = 0; k < P.n_sub; k++) {
            if (kind[k] == 1) {
                const double rx = X.v.x - org[k][0], ry = X.v.y - org[k][1], rz = X.v.z - org[k][2];
                J[0][k] = aw[k][1] * rz - aw[k][2] * ry;
                J[1][k] = aw[k][2] * rx - aw[k][0] * rz;
                J[2][k] = aw[k][0] * ry - aw[k][1] * rx;
                J[3][k] = aw[k][0]; J[4][k] = aw[k][1]; J[5][k] = aw[k][2];
            } else if (kind[k] == 2) {
                J[0][k] = aw[k][0]; J[1][k] = aw[k][1]; J[2][k] = aw[k][2];
                J[3][k] = J[4][k] = J[5][k] = 0.0;
            } else {
                for (int i = 0; i < 6; i++) J[i][k] = 0.0;
            }
        }
        bool shared = false;
        for (int k = 0; k < P.n_sub; k++) shared |= kind[k] < 0;
        if (shared) { /* variables shared by several joints: walk the chain again and add every joint's column */
            Xf<double> Y = base;
            for (int f = 0; f < P.n_fr; f++) {
                const IkFrame &F = fr[f];
                const Xf<double> E = xload(F.E);
                if (F.type == 0) { Y = xmul(Y, E); continue; }
                Vec3<double> ax;
                ax.x = F.ax[0]; ax.y = F.ax[1]; ax.z = F.ax[2];
                Y = fk_joint<double>(E, ax, F.off, F.type, F.col >= 0 ? q[F.col] : F.qfix, &Y);
                if (F.col < 0 || kind[F.col] >= 0) continue;
                const double inv = 1.0 / sqrt(ax.x * ax.x + ax.y * ax.y + ax.z * ax.z);
                const Vec3<double> w = qrot(Y.r, ax * inv);
                const int k = F.col;
                if (F.type == 1) {
                    const double rx = X.v.x - Y.v.x, ry = X.v.y - Y.v.y, rz = X.v.z - Y.v.z;
                    J[0][k] += w.y * rz - w.z * ry;
                    J[1][k] += w.z * rx - w.x * rz;
                    J[2][k] += w.x * ry - w.y * rx;
                    J[3][k] += w.x; J[4][k] += w.y; J[5][k] += w.z;
                } else {
                    J[0][k] += w.x; J[1][k] += w.y; J[2][k] += w.z;
                }
            }
        }
        double A[6][6];
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) {
                double acc = 0;
                for (int k = 0; k < P.n_sub; k++) acc += J[i][k] * J[j][k];
                A[i][j] = acc;
                A[j][i] = acc;
            }
        const double en = ep + eo;
        const double lambda2 = fmin(1e-2, 1e-2 * en * en) + 1e-9; /* damping shrinks as the solution is approached */
        double y[6];
        if (!ik_solve6(A, e, lambda2, y)) { done = true; continue; }
        double big = 0, dq[TMK_IK_MAXCOL];
        for (int k = 0; k < P.n_sub; k++) {
            double acc = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) acc += J[i][k] * y[i];
            dq[k] = acc;
            big = fmax(big, fabs(acc));
        }
        const double scale = big > 0.4 ? 0.4 / big : 1.0;
        for (int k = 0; k < P.n_sub; k++) q[k] = fmin(hi[k], fmax(lo[k], q[k] + scale * dq[k]));
    }
    if (active) {
        for (int k = 0; k < P.n_sub; k++) Q[(size_t)s * P.n_sub + k] = q[k];
        err[2 * s] = ep;
        err[2 * s + 1] = eo;
    }
    if (s == 0) *iters = it;
}

} // namespace tmk
