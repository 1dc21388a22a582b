/* mp.c - one motion query on top of the batched validity check: workspace-goal IK, RRT-Connect,
 * shortcut simplification.  Host C; every state / edge validity decision goes through the C ABI
 * (aa_rx_cl_check_batch, aa_rx_cl_check_edges), i.e. through the CUDA kernels.
 *
 * What it stands in for (all external to the reference tree):
 *   robray:motion-plan  ->  aa_rx_mp_create / set_start / set_wsgoal / set_simplify /
 *   set_track_collisions / aa_rx_mp_plan(timeout)        lisp/python/tmsmtpy.lisp:55-60
 *   OMPL RRTConnect (range = 20 % of the state-space extent), DiscreteMotionValidator
 *   (resolution = 1 % of the extent) and PathSimplifier.
 * The planner is frontier-parallel: every iteration proposes K extensions and checks them, and
 * the K connection attempts that follow, with one aa_rx_cl_check_edges call each.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tmk_internal.h"

#define MP_K 128        /* extensions proposed per iteration */
#define MP_IK_SEEDS 32
#define MP_IK_ITERS 120
#define MP_MAX_GOALS 16
#define MP_IK_POLISH 8
#define MP_MAX_BLOCKED 2048

typedef struct {
    size_t n, cap, dim;
    double *q;   /* n x dim */
    int *parent; /* -1 = root */
} tree;

struct aa_rx_mp {
    const struct aa_rx_sg_sub *ssg;
    const struct aa_rx_sg *sg;
    size_t n_all, n_sub;
    const aa_rx_config_id *ids;
    double *lo, *hi;
    double extent, range, seg_len;
    double *q_start_all;
    int have_start;
    struct aa_rx_cl *cl;
    size_t n_goal;
    double goals[MP_MAX_GOALS][TMK_MP_MAXSUB];
    int simplify, track;
    uint64_t rng;
    size_t n_blocked;
    double *blocked; /* the first invalid state of every rejected edge (for the collision feed-back) */
    struct tmk_mp_stats st;
    int log_on;          /* measurement aid: keep every checked edge so that a CPU baseline can replay them */
    size_t log_n, log_cap;
    double *log_q0, *log_q1;
};

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static uint64_t rng_next(uint64_t *s) { /* splitmix64 */
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double rng_unit(uint64_t *s) { return (double)(rng_next(s) >> 11) * (1.0 / 9007199254740992.0); }

static double dist(const double *a, const double *b, size_t n) {
    double s = 0;
    for (size_t k = 0; k < n; k++) s += (a[k] - b[k]) * (a[k] - b[k]);
    return sqrt(s);
}

static void tree_init(tree *t, size_t dim) { memset(t, 0, sizeof(*t)); t->dim = dim; }
static void tree_free(tree *t) { free(t->q); free(t->parent); }
static int tree_add(tree *t, const double *q, int parent) {
    if (t->n == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 256;
        t->q = (double *)realloc(t->q, sizeof(double) * t->cap * t->dim);
        t->parent = (int *)realloc(t->parent, sizeof(int) * t->cap);
    }
    memcpy(t->q + t->n * t->dim, q, sizeof(double) * t->dim);
    t->parent[t->n] = parent;
    return (int)t->n++;
}
static int tree_nearest(const tree *t, const double *q, double *d_out) {
    int best = 0;
    double bd = INFINITY;
    for (size_t i = 0; i < t->n; i++) {
        double s = 0;
        const double *p = t->q + i * t->dim;
        for (size_t k = 0; k < t->dim; k++) s += (p[k] - q[k]) * (p[k] - q[k]);
        if (s < bd) { bd = s; best = (int)i; }
    }
    if (d_out) *d_out = sqrt(bd);
    return best;
}

/* ------------------------------------------------------------------ create / destroy / options */
struct aa_rx_mp *aa_rx_mp_create(const struct aa_rx_sg_sub *ssg) {
    struct aa_rx_mp *mp = (struct aa_rx_mp *)calloc(1, sizeof(*mp));
    mp->ssg = ssg;
    mp->sg = tmk_sub_sg(ssg);
    mp->n_all = aa_rx_sg_config_count(mp->sg);
    mp->n_sub = aa_rx_sg_sub_config_count(ssg);
    mp->ids = aa_rx_sg_sub_configs(ssg);
    if (mp->n_sub < 1) tmk_die("aa_rx_mp_create: the sub-scene-graph has no configuration variable");
    if (mp->n_sub > TMK_MP_MAXSUB) tmk_die("aa_rx_mp_create: %zu configuration variables > %d", mp->n_sub, TMK_MP_MAXSUB);
    mp->lo = (double *)malloc(sizeof(double) * mp->n_sub);
    mp->hi = (double *)malloc(sizeof(double) * mp->n_sub);
    double e2 = 0;
    for (size_t k = 0; k < mp->n_sub; k++) {
        if (aa_rx_sg_get_limit_pos(mp->sg, mp->ids[k], &mp->lo[k], &mp->hi[k]) != 0) { mp->lo[k] = -M_PI; mp->hi[k] = M_PI; }
        e2 += (mp->hi[k] - mp->lo[k]) * (mp->hi[k] - mp->lo[k]);
    }
    mp->extent = sqrt(e2);
    mp->range = 0.2 * mp->extent;    /* OMPL RRTConnect default range */
    mp->seg_len = 0.01 * mp->extent; /* longestValidSegmentFraction */
    mp->q_start_all = (double *)calloc(mp->n_all + 1, sizeof(double));
    mp->simplify = 0;
    mp->rng = 0x1234567ull;
    mp->blocked = (double *)malloc(sizeof(double) * MP_MAX_BLOCKED * mp->n_sub);
    return mp;
}

void aa_rx_mp_destroy(struct aa_rx_mp *mp) {
    if (!mp) return;
    if (mp->cl) aa_rx_cl_destroy(mp->cl);
    free(mp->lo); free(mp->hi); free(mp->q_start_all); free(mp->blocked);
    free(mp->log_q0); free(mp->log_q1);
    free(mp);
}

void aa_rx_mp_set_simplify(struct aa_rx_mp *mp, int simplify) { mp->simplify = simplify; }
void aa_rx_mp_set_track_collisions(struct aa_rx_mp *mp, int track) { mp->track = track; }
void aa_rx_mp_set_seed(struct aa_rx_mp *mp, uint64_t seed) { mp->rng = seed * 0x9e3779b97f4a7c15ull + 1; }
void aa_rx_mp_stats(const struct aa_rx_mp *mp, struct tmk_mp_stats *out) { *out = mp->st; }
void aa_rx_mp_edge_log(struct aa_rx_mp *mp, int enable) { mp->log_on = enable; mp->log_n = 0; }
size_t aa_rx_mp_edge_log_get(const struct aa_rx_mp *mp, const double **q0, const double **q1) {
    *q0 = mp->log_q0; *q1 = mp->log_q1;
    return mp->log_n;
}

/* start configuration; what collides there is allowed for this query (in-tree
 * analogue demo/tmpexec.c:118) */
void aa_rx_mp_set_start(struct aa_rx_mp *mp, size_t n_all, const double *q_all) {
    if (n_all != mp->n_all) tmk_die("aa_rx_mp_set_start: n_all = %zu, scene graph has %zu", n_all, mp->n_all);
    memcpy(mp->q_start_all, q_all, sizeof(double) * n_all);
    if (mp->cl) aa_rx_cl_destroy(mp->cl);
    mp->cl = aa_rx_cl_create(mp->sg);
    size_t n_f = aa_rx_sg_frame_count(mp->sg);
    double *ab = (double *)malloc(sizeof(double) * 7 * (n_f + 1));
    aa_rx_sg_tf(mp->sg, n_all, q_all, n_f, NULL, 7, ab, 7);
    struct aa_rx_cl_set *set = aa_rx_cl_set_create(mp->sg);
    if (aa_rx_cl_check(mp->cl, n_f, ab, 7, set)) aa_rx_cl_allow_set(mp->cl, set);
    aa_rx_cl_set_destroy(set);
    free(ab);
    mp->have_start = 1;
    mp->n_goal = 0;
    memset(&mp->st, 0, sizeof(mp->st));
}

static int in_limits(const struct aa_rx_mp *mp, const double *q) {
    for (size_t k = 0; k < mp->n_sub; k++)
        if (q[k] < mp->lo[k] - 1e-9 || q[k] > mp->hi[k] + 1e-9) return 0;
    return 1;
}

static int config_valid(struct aa_rx_mp *mp, const double *q_sub) {
    uint32_t bits[2] = {0, 0};
    aa_rx_cl_check_batch(mp->cl, 1, mp->n_sub, mp->ids, mp->n_all, mp->q_start_all, q_sub, mp->n_sub, bits);
    return (int)(bits[0] & 1u);
}

int aa_rx_mp_set_goal(struct aa_rx_mp *mp, size_t n_sub, const double *q_sub) {
    if (!mp->have_start) tmk_die("aa_rx_mp_set_goal: set the start first");
    if (n_sub != mp->n_sub) tmk_die("aa_rx_mp_set_goal: n_sub mismatch");
    if (!in_limits(mp, q_sub) || !config_valid(mp, q_sub)) return 0;
    if (mp->n_goal < MP_MAX_GOALS) memcpy(mp->goals[mp->n_goal++], q_sub, sizeof(double) * n_sub);
    mp->st.n_goals = mp->n_goal;
    return 1;
}

/* ------------------------------------------------------------------ inverse kinematics */
/* damped least squares over the chain, evaluated on the device for all seeds at once: ik.cuh */
int aa_rx_mp_set_wsgoal(struct aa_rx_mp *mp, const double *E7) {
    if (!mp->have_start) tmk_die("aa_rx_mp_set_wsgoal: set the start first");
    double t0 = now_s();
    double goal[7];
    memcpy(goal, E7, sizeof(goal));
    double gn = sqrt(goal[0] * goal[0] + goal[1] * goal[1] + goal[2] * goal[2] + goal[3] * goal[3]);
    for (int k = 0; k < 4; k++) goal[k] /= gn;
    const aa_rx_frame_id *frames;
    size_t n_fr = tmk_sub_frames(mp->ssg, &frames);
    const size_t S = MP_IK_SEEDS, n = mp->n_sub;
    double *Q = (double *)malloc(sizeof(double) * S * n);
    for (size_t s = 0; s < S; s++)
        for (size_t k = 0; k < n; k++) {
            double q0 = mp->q_start_all[mp->ids[k]];
            if (s == 0) Q[s * n + k] = q0;
            else if (s < 8) { /* near the start: short motions are preferred */
                double w = 0.15 * (double)s;
                double v = q0 + w * (rng_unit(&mp->rng) - 0.5) * (mp->hi[k] - mp->lo[k]);
                Q[s * n + k] = fmin(mp->hi[k], fmax(mp->lo[k], v));
            } else Q[s * n + k] = mp->lo[k] + rng_unit(&mp->rng) * (mp->hi[k] - mp->lo[k]);
        }
    /* all seeds iterated to convergence in one launch (double precision, ik.cuh) */
    double *err = (double *)malloc(sizeof(double) * 2 * S);
    const double tol = 1e-9;
    mp->st.n_ik_iterations += (size_t)tmk_dev_ik(mp->sg, n_fr, frames, n, mp->ids, mp->q_start_all, mp->lo, mp->hi, goal, S, Q,
                                                 err, MP_IK_ITERS, 8, 16, tol);
    double cand[MP_IK_SEEDS][TMK_MP_MAXSUB];
    double cand_cost[MP_IK_SEEDS];
    size_t n_cand = 0;
    /* nearest to the start first; the MP_IK_POLISH most promising converged candidates are kept */
    size_t order[MP_IK_SEEDS], n_ord = 0;
    double ocost[MP_IK_SEEDS];
    double start_sub[TMK_MP_MAXSUB];
    for (size_t k = 0; k < n; k++) start_sub[k] = mp->q_start_all[mp->ids[k]];
    for (size_t s = 0; s < S; s++) {
        if (!(err[2 * s] < tol && err[2 * s + 1] < tol)) continue;
        double c = dist(Q + s * n, start_sub, n);
        size_t pos = n_ord++;
        while (pos > 0 && ocost[pos - 1] > c) { order[pos] = order[pos - 1]; ocost[pos] = ocost[pos - 1]; pos--; }
        order[pos] = s; ocost[pos] = c;
    }
    for (size_t oi = 0; oi < n_ord && n_cand < MP_IK_POLISH; oi++) {
        const double *q = Q + order[oi] * n;
        if (!in_limits(mp, q)) continue;
        int dup = 0;
        for (size_t c = 0; c < n_cand; c++) dup |= dist(cand[c], q, n) < 1e-3;
        if (dup) continue;
        memcpy(cand[n_cand], q, sizeof(double) * n);
        cand_cost[n_cand] = ocost[oi];
        n_cand++;
    }
    /* collision-free candidates, nearest to the start first */
    mp->n_goal = 0;
    if (n_cand) {
        uint32_t bits[(MP_IK_SEEDS + 31) / 32 + 1];
        aa_rx_cl_check_batch(mp->cl, n_cand, n, mp->ids, mp->n_all, mp->q_start_all, &cand[0][0], TMK_MP_MAXSUB, bits);
        for (;;) {
            int best = -1;
            for (size_t c = 0; c < n_cand; c++)
                if (((bits[c >> 5] >> (c & 31)) & 1u) && (best < 0 || cand_cost[c] < cand_cost[best])) best = (int)c;
            if (best < 0 || mp->n_goal == MP_MAX_GOALS) break;
            memcpy(mp->goals[mp->n_goal++], cand[best], sizeof(double) * n);
            bits[best >> 5] &= ~(1u << (best & 31));
        }
    }
    free(Q); free(err);
    mp->st.n_goals = mp->n_goal;
    mp->st.ik_s += now_s() - t0;
    return (int)mp->n_goal;
}

/* ------------------------------------------------------------------ RRT-Connect, frontier-parallel */
static void remember_blocked(struct aa_rx_mp *mp, const double *q) {
    if (mp->n_blocked < MP_MAX_BLOCKED) memcpy(mp->blocked + (mp->n_blocked++) * mp->n_sub, q, sizeof(double) * mp->n_sub);
}

/* k-th (0-based) interior index OMPL's checkMotion tests on an edge of nd segments: FIFO bisection of 1 .. nd-1
 * (the device restatement is ompl_mid in dev_math.cuh) */
static int ompl_mid_host(int nd, int k) {
    int *qa = (int *)malloc(sizeof(int) * 2 * (size_t)(nd + 2)), head = 0, tail = 0, mid = 1;
    qa[tail++] = 1; qa[tail++] = nd - 1;
    for (int it = 0; it <= k && head < tail; it++) {
        const int a = qa[head++], b = qa[head++];
        mid = (a + b) / 2;
        if (a < mid) { qa[tail++] = a; qa[tail++] = mid - 1; }
        if (mid < b) { qa[tail++] = mid + 1; qa[tail++] = b; }
    }
    free(qa);
    return mid;
}

/* the state at position `pos` of OMPL's test order on the edge a -> b (0 = the end state b) */
static void edge_state(const struct aa_rx_mp *mp, const double *a, const double *b, int pos, double *out) {
    const size_t dim = mp->n_sub;
    if (pos <= 0) { memcpy(out, b, sizeof(double) * dim); return; }
    double s = 0;
    for (size_t k = 0; k < dim; k++) s += (b[k] - a[k]) * (b[k] - a[k]);
    int nd = (int)ceil(sqrt(s) / mp->seg_len);
    if (nd < 1) nd = 1;
    const double t = (double)ompl_mid_host(nd, pos - 1) / (double)nd;
    for (size_t k = 0; k < dim; k++) out[k] = a[k] + (b[k] - a[k]) * t;
}

static size_t check_edges(struct aa_rx_mp *mp, size_t n_edge, const double *Q0, const double *Q1, uint32_t *bits) {
    mp->st.n_edge_calls++;
    mp->st.n_edges_checked += n_edge;
    if (mp->log_on) {
        if (mp->log_n + n_edge > mp->log_cap) {
            mp->log_cap = 2 * (mp->log_n + n_edge) + 1024;
            mp->log_q0 = (double *)realloc(mp->log_q0, sizeof(double) * mp->log_cap * mp->n_sub);
            mp->log_q1 = (double *)realloc(mp->log_q1, sizeof(double) * mp->log_cap * mp->n_sub);
        }
        memcpy(mp->log_q0 + mp->log_n * mp->n_sub, Q0, sizeof(double) * n_edge * mp->n_sub);
        memcpy(mp->log_q1 + mp->log_n * mp->n_sub, Q1, sizeof(double) * n_edge * mp->n_sub);
        mp->log_n += n_edge;
    }
    if (!mp->track || mp->n_blocked >= MP_MAX_BLOCKED)
        return aa_rx_cl_check_edges(mp->cl, n_edge, mp->n_sub, mp->ids, mp->n_all, mp->q_start_all, Q0, Q1, mp->n_sub,
                                    mp->seg_len, bits, NULL);
    /* :track-collisions (lisp/python/tmsmtpy.lisp:59): remember the state that actually blocked every rejected
     * edge - the first invalid one in OMPL's test order, which may be an interior state of an edge whose end point
     * is free - so that aa_rx_mp_collisions can report the frame pairs colliding there */
    int32_t *first = (int32_t *)malloc(sizeof(int32_t) * (n_edge + 1));
    const size_t bad = aa_rx_cl_check_edges(mp->cl, n_edge, mp->n_sub, mp->ids, mp->n_all, mp->q_start_all, Q0, Q1, mp->n_sub,
                                            mp->seg_len, bits, first);
    double q[TMK_MP_MAXSUB];
    for (size_t e = 0; e < n_edge && bad; e++)
        if (first[e] >= 0) {
            edge_state(mp, Q0 + e * mp->n_sub, Q1 + e * mp->n_sub, first[e], q);
            remember_blocked(mp, q);
        }
    free(first);
    return bad;
}

/* Path simplification in the order of OMPL's PathSimplifier::simplify (what :simplify t asks for): first
 * reduceVertices - drop vertices whose neighbours can be joined directly -, then shortcutPath - join points ON
 * the path's segments -, then a final check of the whole path.  Both stages are batched: where OMPL tries one
 * random shortcut at a time, every candidate of a round goes through one aa_rx_cl_check_edges call.
 *
 * reduce_vertices: the shortest vertex sub-sequence through valid straight edges (the fixed point reduceVertices
 * converges to).  All vertex pairs when they fit one call; for long raw paths only pairs at most `window` vertices
 * apart, repeated until nothing changes (OMPL's rangeRatio plays the same role).  path: n x dim, returns the new n. */
#define MP_SHORTCUT_EDGES 8192
static size_t reduce_vertices(struct aa_rx_mp *mp, double *path, size_t n) {
    const size_t dim = mp->n_sub;
    for (int round = 0; round < 8 && n >= 3; round++) {
        size_t window = n;
        while (window > 2 && n * window > MP_SHORTCUT_EDGES) window = window * 3 / 4;
        size_t n_e = 0, cap_e = n * window + 8;
        double *Q0 = (double *)malloc(sizeof(double) * dim * cap_e), *Q1 = (double *)malloc(sizeof(double) * dim * cap_e);
        int *ei = (int *)malloc(sizeof(int) * cap_e), *ej = (int *)malloc(sizeof(int) * cap_e);
        for (size_t i = 0; i < n; i++)
            for (size_t j = i + 2; j < n && j <= i + window; j++) {
                memcpy(Q0 + n_e * dim, path + i * dim, sizeof(double) * dim);
                memcpy(Q1 + n_e * dim, path + j * dim, sizeof(double) * dim);
                ei[n_e] = (int)i; ej[n_e] = (int)j;
                n_e++;
            }
        uint32_t *bits = (uint32_t *)calloc(n_e / 32 + 2, sizeof(uint32_t));
        if (n_e) check_edges(mp, n_e, Q0, Q1, bits);
        /* DP over the DAG: consecutive edges are known valid; candidate edges are sorted by i, scanned by j */
        double *cost = (double *)malloc(sizeof(double) * n);
        int *prev = (int *)malloc(sizeof(int) * n);
        for (size_t j = 0; j < n; j++) { cost[j] = INFINITY; prev[j] = -1; }
        cost[0] = 0;
        for (size_t j = 1; j < n; j++) {
            double c = cost[j - 1] + dist(path + (j - 1) * dim, path + j * dim, dim);
            if (c < cost[j]) { cost[j] = c; prev[j] = (int)j - 1; }
        }
        /* edges are generated in increasing i, so cost[i] is final when edge (i, j) relaxes j (Dijkstra order on a DAG) */
        for (size_t e = 0; e < n_e; e++)
            if ((bits[e >> 5] >> (e & 31)) & 1u) {
                const size_t i = (size_t)ei[e], j = (size_t)ej[e];
                const double c2 = cost[i] + dist(path + i * dim, path + j * dim, dim);
                if (c2 < cost[j]) {
                    cost[j] = c2; prev[j] = (int)i;
                    for (size_t k = j + 1; k < n; k++) { /* the chain of consecutive edges behind j */
                        const double c = cost[k - 1] + dist(path + (k - 1) * dim, path + k * dim, dim);
                        if (c < cost[k]) { cost[k] = c; prev[k] = (int)k - 1; } else break;
                    }
                }
            }
        int *keep = (int *)malloc(sizeof(int) * n);
        size_t m = 0;
        for (int v = (int)n - 1; v >= 0; v = prev[v]) { keep[m++] = v; if (v == 0) break; }
        double *out = (double *)malloc(sizeof(double) * n * dim);
        for (size_t k = 0; k < m; k++) memcpy(out + k * dim, path + (size_t)keep[m - 1 - k] * dim, sizeof(double) * dim);
        memcpy(path, out, sizeof(double) * m * dim);
        free(out); free(keep); free(cost); free(prev); free(bits); free(Q0); free(Q1); free(ei); free(ej);
        const int done = m == n || window >= n; /* all pairs were tried, or nothing moved */
        n = m;
        if (done) break;
    }
    return n;
}

/* point at arc length d along the polyline (cum[i] = length up to vertex i); returns the segment index */
static size_t point_at(const double *path, const double *cum, size_t n, size_t dim, double d, double *out) {
    size_t i = 0;
    while (i + 2 < n && cum[i + 1] <= d) i++;
    const double len = cum[i + 1] - cum[i];
    const double t = len > 0 ? fmin(1.0, fmax(0.0, (d - cum[i]) / len)) : 0.0;
    for (size_t k = 0; k < dim; k++) out[k] = path[i * dim + k] + t * (path[(i + 1) * dim + k] - path[i * dim + k]);
    return i;
}

/* shortcut_path: MP_SC random pairs of points on the path per round (the second within a third of the path length
 * of the first, OMPL's rangeRatio 0.33).  A shortcut P0 -> P1 replaces the path between a point of segment i and a
 * point of segment j > i; the two stubs it leaves (vertex i -> P0, P1 -> vertex j+1) are sub-segments of validated
 * edges, but a discretised check of a sub-segment samples other states than the check of the segment did - so all
 * three edges of every candidate go into the round's ONE batched check and a candidate counts only if all three are
 * valid (OMPL validates the repaired path afterwards, checkAndRepair; here nothing invalid is ever inserted).  The valid
 * candidates are applied by decreasing saving as long as they do not overlap and keep a vertex between them.
 * *path may be reallocated (every applied shortcut adds up to two vertices).  Returns the new n. */
#define MP_SC 96
static size_t shortcut_path(struct aa_rx_mp *mp, double **path_io, size_t n) {
    const size_t dim = mp->n_sub;
    double *E0 = (double *)malloc(sizeof(double) * 3 * MP_SC * dim), *E1 = (double *)malloc(sizeof(double) * 3 * MP_SC * dim);
    double d0[MP_SC], save[MP_SC];
    size_t s0[MP_SC], s1[MP_SC];
    uint32_t bits[3 * MP_SC / 32 + 2];
    for (int round = 0; round < 2 && n >= 3; round++) {
        double *path = *path_io;
        double *cum = (double *)malloc(sizeof(double) * n);
        cum[0] = 0;
        for (size_t i = 1; i < n; i++) cum[i] = cum[i - 1] + dist(path + (i - 1) * dim, path + i * dim, dim);
        const double L = cum[n - 1];
        if (!(L > 0)) { free(cum); break; }
        int n_c = 0;
        for (int k = 0; k < MP_SC; k++) {
            double a = rng_unit(&mp->rng) * L, b = a + (0.02 + 0.31 * rng_unit(&mp->rng)) * L;
            if (b > L) b = L;
            double *P0 = E0 + (size_t)(3 * n_c) * dim, *P1 = E1 + (size_t)(3 * n_c) * dim;
            const size_t sa = point_at(path, cum, n, dim, a, P0), sb = point_at(path, cum, n, dim, b, P1);
            if (sa == sb) continue; /* both on one segment: nothing to gain */
            const double direct = dist(P0, P1, dim);
            if (b - a - direct < 1e-4 * L) continue;
            /* edges 3k+1, 3k+2: the stubs vertex sa -> P0 and P1 -> vertex sb+1 */
            memcpy(E0 + (size_t)(3 * n_c + 1) * dim, path + sa * dim, sizeof(double) * dim);
            memcpy(E1 + (size_t)(3 * n_c + 1) * dim, P0, sizeof(double) * dim);
            memcpy(E0 + (size_t)(3 * n_c + 2) * dim, P1, sizeof(double) * dim);
            memcpy(E1 + (size_t)(3 * n_c + 2) * dim, path + (sb + 1) * dim, sizeof(double) * dim);
            d0[n_c] = a; s0[n_c] = sa; s1[n_c] = sb; save[n_c] = b - a - direct;
            n_c++;
        }
        double saved = 0;
        if (n_c) {
            check_edges(mp, (size_t)(3 * n_c), E0, E1, bits);
            /* greedy: largest saving first; intervals must not overlap and must keep a vertex between them */
            int order[MP_SC], n_ok = 0, take[MP_SC], n_take = 0;
            for (int k = 0; k < n_c; k++) {
                const uint32_t e = 3u * (uint32_t)k;
                if (((bits[e >> 5] >> (e & 31)) & 1u) && ((bits[(e + 1) >> 5] >> ((e + 1) & 31)) & 1u) && ((bits[(e + 2) >> 5] >> ((e + 2) & 31)) & 1u))
                    order[n_ok++] = k;
            }
            for (int x = 1; x < n_ok; x++) { /* insertion sort by saving, descending */
                int v = order[x], y = x;
                while (y > 0 && save[order[y - 1]] < save[v]) { order[y] = order[y - 1]; y--; }
                order[y] = v;
            }
            for (int x = 0; x < n_ok; x++) {
                const int k = order[x];
                int clash = 0;
                for (int y = 0; y < n_take; y++) clash |= !(s1[k] < s0[take[y]] || s0[k] > s1[take[y]]);
                if (!clash) take[n_take++] = k;
            }
            if (n_take) {
                for (int x = 1; x < n_take; x++) { /* by position along the path */
                    int v = take[x], y = x;
                    while (y > 0 && d0[take[y - 1]] > d0[v]) { take[y] = take[y - 1]; y--; }
                    take[y] = v;
                }
                double *out = (double *)malloc(sizeof(double) * dim * (n + 2 * (size_t)n_take + 2));
                size_t m = 0, i = 0;
                for (int x = 0; x < n_take; x++) {
                    const int k = take[x];
                    while (i <= s0[k]) { memcpy(out + m * dim, path + i * dim, sizeof(double) * dim); m++; i++; }
                    memcpy(out + m * dim, E0 + (size_t)(3 * k) * dim, sizeof(double) * dim); m++;
                    memcpy(out + m * dim, E1 + (size_t)(3 * k) * dim, sizeof(double) * dim); m++;
                    i = s1[k] + 1; /* the vertices the shortcut bypasses */
                    saved += save[k];
                }
                while (i < n) { memcpy(out + m * dim, path + i * dim, sizeof(double) * dim); m++; i++; }
                /* a shortcut point that coincides with a kept vertex would leave a zero-length edge */
                size_t w = 1;
                for (size_t r = 1; r < m; r++) {
                    if (dist(out + (w - 1) * dim, out + r * dim, dim) > 1e-12) {
                        if (w != r) memmove(out + w * dim, out + r * dim, sizeof(double) * dim);
                        w++;
                    } else if (r + 1 == m && w > 1) /* the goal itself: keep it bit for bit */
                        memmove(out + (w - 1) * dim, out + r * dim, sizeof(double) * dim);
                }
                free(*path_io);
                *path_io = out;
                n = w;
            }
        }
        free(cum);
        if (saved < 0.02 * L) break; /* a further round is a GPU round trip for a few percent */
    }
    free(E0); free(E1);
    return n;
}

/* reduceVertices -> shortcutPath, OMPL's order */
static size_t simplify_path(struct aa_rx_mp *mp, double **path_io, size_t n) {
    n = reduce_vertices(mp, *path_io, n);
    if (n < 3) return n;
    return shortcut_path(mp, path_io, n);
}

int aa_rx_mp_plan(struct aa_rx_mp *mp, double timeout, size_t *n_path, double **path_all) {
    *n_path = 0;
    *path_all = NULL;
    if (!mp->have_start) tmk_die("aa_rx_mp_plan: set the start first");
    if (mp->n_goal == 0) return -1;
    const size_t dim = mp->n_sub;
    double t0 = now_s();
    mp->n_blocked = 0;
    double start[TMK_MP_MAXSUB];
    for (size_t k = 0; k < dim; k++) start[k] = mp->q_start_all[mp->ids[k]];

    tree T[2]; /* T[0] grows from the start, T[1] from the goals */
    tree_init(&T[0], dim);
    tree_init(&T[1], dim);
    tree_add(&T[0], start, -1);
    for (size_t g = 0; g < mp->n_goal; g++) tree_add(&T[1], mp->goals[g], -1);

    double *Q0 = (double *)malloc(sizeof(double) * MP_K * dim), *Q1 = (double *)malloc(sizeof(double) * MP_K * dim);
    int near_idx[MP_K], new_idx[MP_K];
    uint32_t bits[MP_K / 32 + 2];
    int found_a = -1, found_b = -1; /* nodes of T[0] / T[1] joined by a valid straight segment */

    /* the goals may be directly reachable */
    {
        size_t n_e = mp->n_goal;
        for (size_t g = 0; g < n_e; g++) { memcpy(Q0 + g * dim, start, sizeof(double) * dim); memcpy(Q1 + g * dim, mp->goals[g], sizeof(double) * dim); }
        check_edges(mp, n_e, Q0, Q1, bits);
        for (size_t g = 0; g < n_e && found_a < 0; g++)
            if ((bits[g >> 5] >> (g & 31)) & 1u) { found_a = 0; found_b = (int)g; }
    }

    int a = 0; /* tree being extended */
    while (found_a < 0 && now_s() - t0 < timeout) {
        mp->st.n_iterations++;
        tree *Ta = &T[a], *Tb = &T[1 - a];
        /* K extensions of Ta towards uniform samples */
        for (int k = 0; k < MP_K; k++) {
            double s[TMK_MP_MAXSUB], d;
            for (size_t j = 0; j < dim; j++) s[j] = mp->lo[j] + rng_unit(&mp->rng) * (mp->hi[j] - mp->lo[j]);
            near_idx[k] = tree_nearest(Ta, s, &d);
            const double *nq = Ta->q + (size_t)near_idx[k] * dim;
            double f = d > mp->range ? mp->range / d : 1.0;
            for (size_t j = 0; j < dim; j++) {
                Q0[k * dim + j] = nq[j];
                Q1[k * dim + j] = nq[j] + f * (s[j] - nq[j]);
            }
        }
        check_edges(mp, MP_K, Q0, Q1, bits);
        int n_new = 0;
        for (int k = 0; k < MP_K; k++) {
            if ((bits[k >> 5] >> (k & 31)) & 1u) new_idx[n_new++] = tree_add(Ta, Q1 + k * dim, near_idx[k]);
        }
        if (n_new) {
            /* connect: straight segments from the nearest node of the other tree to every new node */
            for (int k = 0; k < n_new; k++) {
                const double *nq = Ta->q + (size_t)new_idx[k] * dim;
                near_idx[k] = tree_nearest(Tb, nq, NULL);
                memcpy(Q0 + k * dim, Tb->q + (size_t)near_idx[k] * dim, sizeof(double) * dim);
                memcpy(Q1 + k * dim, nq, sizeof(double) * dim);
            }
            check_edges(mp, (size_t)n_new, Q0, Q1, bits);
            double best = INFINITY;
            for (int k = 0; k < n_new; k++)
                if ((bits[k >> 5] >> (k & 31)) & 1u) {
                    double d = dist(Q0 + k * dim, Q1 + k * dim, dim);
                    if (d < best) {
                        best = d;
                        if (a == 0) { found_a = new_idx[k]; found_b = near_idx[k]; }
                        else { found_a = near_idx[k]; found_b = new_idx[k]; }
                    }
                }
        }
        a = 1 - a;
    }
    mp->st.n_nodes = T[0].n + T[1].n;
    mp->st.rrt_s += now_s() - t0;
    int rc = -1;
    if (found_a >= 0) {
        /* start -> found_a (reverse of the parent walk), then found_b -> its goal root */
        size_t na = 0, nb = 0;
        for (int v = found_a; v >= 0; v = T[0].parent[v]) na++;
        for (int v = found_b; v >= 0; v = T[1].parent[v]) nb++;
        size_t n = na + nb;
        double *path = (double *)malloc(sizeof(double) * (n + 1) * dim);
        size_t k = na;
        for (int v = found_a; v >= 0; v = T[0].parent[v]) memcpy(path + (--k) * dim, T[0].q + (size_t)v * dim, sizeof(double) * dim);
        k = na;
        for (int v = found_b; v >= 0; v = T[1].parent[v]) memcpy(path + (k++) * dim, T[1].q + (size_t)v * dim, sizeof(double) * dim);
        if (mp->simplify) {
            double ts = now_s();
            const int track = mp->track;
            mp->track = 0; /* rejected shortcuts are not obstacles between start and goal */
            n = simplify_path(mp, &path, n);
            mp->track = track;
            mp->st.simplify_s += now_s() - ts;
        }
        double *out = (double *)malloc(sizeof(double) * n * mp->n_all);
        for (size_t i = 0; i < n; i++) {
            memcpy(out + i * mp->n_all, mp->q_start_all, sizeof(double) * mp->n_all);
            aa_rx_sg_config_set(mp->sg, mp->n_all, dim, mp->ids, path + i * dim, out + i * mp->n_all);
        }
        free(path);
        *n_path = n;
        *path_all = out;
        rc = 0;
    }
    tree_free(&T[0]); tree_free(&T[1]);
    free(Q0); free(Q1);
    return rc;
}

/* frame pairs that blocked extensions and connections: full-report check of the first invalid state of every
 * rejected edge (what the :collision constraint mode consumes, lisp/itmp-rec.lisp:230-247) */
void aa_rx_mp_collisions(struct aa_rx_mp *mp, struct aa_rx_cl_set *set) {
    if (!mp->cl || mp->n_blocked == 0) return;
    uint32_t *bits = (uint32_t *)calloc(mp->n_blocked / 32 + 2, sizeof(uint32_t));
    aa_rx_cl_track_collisions(mp->cl, 1);
    aa_rx_cl_check_batch(mp->cl, mp->n_blocked, mp->n_sub, mp->ids, mp->n_all, mp->q_start_all, mp->blocked, mp->n_sub, bits);
    aa_rx_cl_batch_collisions(mp->cl, set);
    aa_rx_cl_track_collisions(mp->cl, 0);
    free(bits);
}
