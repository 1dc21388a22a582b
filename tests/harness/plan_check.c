/* plan_check.c - a plan executor / validator in plain C on top of libtmkcl.so: the role demo/tmpexec.c
 * plays in the reference (load a scene plugin, read a plan file, validate every state of every motion
 * plan with aa_rx_sg_tf + aa_rx_cl_check, apply the reparent ops between motions).  TEST
 * INFRASTRUCTURE: it shows that a C caller written against <amino/rx/...> call shapes runs against
 * include/tmkit_b200/amino_rx.h unchanged in style; it is not part of the library.
 *
 *   plan_check <plugin.so> <scene name> <start.tmp> <plan.tmp> [resolution fraction, default 0.01]
 *
 * Prints one line per plan operation and a summary; exit status 0 iff no state of any motion collides
 * (after the start-state allowance, demo/tmpexec.c:118).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tmkit_b200/amino_rx.h"
#include "tmkit_b200/tmplan.h" /* the reference's plan-file ABI: include/tmsmt/tmplan.h */

/* the per-state check of the reference executor (tmpexec.c:362-428), one state */
static int check_state(const struct aa_rx_sg *sg, struct aa_rx_cl *cl, struct aa_rx_cl_set *cl_set, const double *q, int verbose) {
    size_t n_q = aa_rx_sg_config_count(sg), n_f = aa_rx_sg_frame_count(sg);
    double *TF_rel = (double *)malloc(sizeof(double) * 7 * n_f), *TF_abs = (double *)malloc(sizeof(double) * 7 * n_f);
    aa_rx_cl_set_clear(cl_set);
    aa_rx_sg_tf(sg, n_q, q, n_f, TF_rel, 7, TF_abs, 7);
    int collision = aa_rx_cl_check(cl, n_f, TF_abs, 7, cl_set);
    if (collision && verbose)
        for (size_t i = 0; i < n_f; i++)
            for (size_t j = 0; j < i; j++)
                if (aa_rx_cl_set_get(cl_set, (int)i, (int)j))
                    fprintf(stderr, "WARNING: collision between %s(%d) <-> %s(%d)\n", aa_rx_sg_frame_name(sg, (int)i), (int)i,
                            aa_rx_sg_frame_name(sg, (int)j), (int)j);
    free(TF_rel); free(TF_abs);
    return collision;
}

/* executor context: what tmpexec.c keeps in struct exec_cx (scene graph and configuration at the current end of
 * the plan) plus the counters this validator prints */
struct exec_cx {
    struct aa_rx_sg *sg;
    double *q;
    size_t n_q;
    double resolution;
    size_t n_states, n_bad, n_motion, n_reparent;
};

/* one operator of the plan, called back by tmplan_map_ops as in tmpexec.c:125, 232-350 */
static void op_function(void *cx_, const struct tmplan_op *op) {
    struct exec_cx *cx = (struct exec_cx *)cx_;
    struct aa_rx_sg *sg = cx->sg;
    size_t n_q = cx->n_q;
    double *q = cx->q;
    switch (tmplan_op_type(op)) {
    case TMPLAN_OP_ACTION:
        printf("action   %s\n", tmplan_op_action_get((const struct tmplan_op_action *)op));
        break;
    case TMPLAN_OP_MOTION_PLAN: {
        struct tmplan_op_motion_plan *top = (struct tmplan_op_motion_plan *)op;
        const double *pts = tmplan_op_motion_plan_path(top);             /* tmpexec.c:265-267 */
        size_t n = tmplan_op_motion_plan_config_count(top);
        size_t n_elt = tmplan_op_motion_plan_path_size(top);
        size_t np = n ? n_elt / n : 0;
        const char **names = (const char **)malloc(sizeof(char *) * (n + 1));
        aa_rx_config_id *ids = (aa_rx_config_id *)malloc(sizeof(int) * (n + 1));
        tmplan_op_motion_plan_vars(top, n, names);                       /* tmpexec.c:271 */
        aa_rx_sg_config_indices(sg, n, names, ids);                      /* map by NAME: tmpexec.c:272 */
        /* the scene allows what collides at the start of the motion (tmpexec.c:118) */
        aa_rx_sg_allow_config(sg, n_q, q);
        struct aa_rx_cl *cl = aa_rx_cl_create(sg);
        struct aa_rx_cl_set *cl_set = aa_rx_cl_set_create(sg);
        /* every waypoint with the reference's single-state check ... */
        size_t bad = 0;
        double *qq = (double *)malloc(sizeof(double) * n_q);
        memcpy(qq, q, sizeof(double) * n_q);
        for (size_t i = 0; i < np; i++) {
            aa_rx_sg_config_set(sg, n_q, n, ids, pts + i * n, qq);       /* tmpexec.c:289-291 */
            bad += check_state(sg, cl, cl_set, qq, 1) != 0;
            cx->n_states++;
        }
        /* ... and every segment at the planner's resolution through the swept-edge entry point */
        if (np > 1 && cx->resolution > 0) {
            double ext2 = 0;
            for (size_t c = 0; c < n; c++) {
                double lo = -M_PI, hi = M_PI;
                aa_rx_sg_get_limit_pos(sg, ids[c], &lo, &hi);
                ext2 += (hi - lo) * (hi - lo);
            }
            uint32_t *bits = (uint32_t *)calloc((np + 31) / 32 + 1, sizeof(uint32_t));
            bad += aa_rx_cl_check_edges(cl, np - 1, n, ids, n_q, q, pts, pts + n, n, cx->resolution * sqrt(ext2), bits, NULL);
            cx->n_states += np - 1;
            free(bits);
        }
        printf("motion   %zu waypoints over %zu joints: %zu colliding states\n", np, n, bad);
        cx->n_bad += bad;
        cx->n_motion++;
        if (np) aa_rx_sg_config_set(sg, n_q, n, ids, pts + (np - 1) * n, q);
        free(qq); free(ids); free(names);
        aa_rx_cl_set_destroy(cl_set);
        aa_rx_cl_destroy(cl);
    } break;
    case TMPLAN_OP_REPARENT: {
        /* reparent at the current configuration, world pose preserved (tmpexec.c:307-348) */
        struct tmplan_op_reparent *top = (struct tmplan_op_reparent *)op;
        const char *frame = tmplan_op_reparent_get_frame(top), *parent = tmplan_op_reparent_get_new_parent(top);
        size_t n_f = aa_rx_sg_frame_count(sg);
        double *TF_abs = (double *)malloc(sizeof(double) * 7 * n_f), rel[7];
        aa_rx_sg_tf(sg, n_q, q, n_f, NULL, 7, TF_abs, 7);
        aa_rx_sg_rel_tf(sg, aa_rx_sg_frame_id(sg, parent), aa_rx_sg_frame_id(sg, frame), TF_abs, 7, rel);
        struct aa_rx_sg *sg2 = aa_rx_sg_copy(sg);
        aa_rx_sg_reparent_name(sg2, parent, frame, rel);
        aa_rx_sg_init(sg2);
        aa_rx_sg_cl_init(sg2);
        aa_rx_sg_destroy(sg);
        cx->sg = sg2;
        free(TF_abs);
        printf("reparent %s -> %s\n", frame, parent);
        cx->n_reparent++;
    } break;
    }
}

/* the first motion plan of the start file sets the start configuration (-q q0.tmp, lisp/driver.lisp:92-95) */
static void start_function(void *cx_, const struct tmplan_op *op) {
    struct exec_cx *cx = (struct exec_cx *)cx_;
    if (tmplan_op_type(op) != TMPLAN_OP_MOTION_PLAN || cx->n_motion) return;
    struct tmplan_op_motion_plan *top = (struct tmplan_op_motion_plan *)op;
    size_t n = tmplan_op_motion_plan_config_count(top), n_elt = tmplan_op_motion_plan_path_size(top);
    if (!n || n_elt < n) return;
    const char **names = (const char **)malloc(sizeof(char *) * n);
    aa_rx_config_id *ids = (aa_rx_config_id *)malloc(sizeof(int) * n);
    tmplan_op_motion_plan_vars(top, n, names);
    aa_rx_sg_config_indices(cx->sg, n, names, ids);
    aa_rx_sg_config_set(cx->sg, cx->n_q, n, ids, tmplan_op_motion_plan_path(top) + (n_elt / n - 1) * n, cx->q);
    cx->n_motion = 1;
    free(names); free(ids);
}

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s plugin.so scene start.tmp plan.tmp [resolution]\n", argv[0]);
        return 2;
    }
    struct exec_cx cx;
    memset(&cx, 0, sizeof(cx));
    cx.resolution = argc > 5 ? atof(argv[5]) : 0.01; /* OMPL longestValidSegmentFraction */
    cx.sg = aa_rx_dl_sg(argv[1], argv[2], NULL);     /* tmpexec.c:172 */
    if (!cx.sg) return 2;
    aa_rx_sg_init(cx.sg);                            /* tmpexec.c:105-106 */
    aa_rx_sg_cl_init(cx.sg);
    cx.n_q = aa_rx_sg_config_count(cx.sg);
    cx.q = (double *)calloc(cx.n_q, sizeof(double));

    struct tmplan *start = tmplan_parse_filename(argv[3], NULL); /* tmpexec.c:111 (region argument unused here) */
    if (!start || tmplan_op_count(start) < 1) return 2;
    tmplan_map_ops(start, start_function, &cx);
    if (!cx.n_motion) return 2;
    cx.n_motion = 0;
    struct tmplan *plan = tmplan_parse_filename(argv[4], NULL);
    if (!plan) return 2;
    tmplan_map_ops(plan, op_function, &cx);          /* tmpexec.c:125 */

    /* final world poses of the blocks, for the caller to compare */
    {
        size_t n_f = aa_rx_sg_frame_count(cx.sg);
        double *TF_abs = (double *)malloc(sizeof(double) * 7 * n_f);
        aa_rx_sg_tf(cx.sg, cx.n_q, cx.q, n_f, NULL, 7, TF_abs, 7);
        for (size_t i = 0; i < n_f; i++)
            if (0 == strncmp(aa_rx_sg_frame_name(cx.sg, (int)i), "block_", 6))
                printf("final    %s %.9f %.9f %.9f\n", aa_rx_sg_frame_name(cx.sg, (int)i), TF_abs[7 * i + 4], TF_abs[7 * i + 5], TF_abs[7 * i + 6]);
        free(TF_abs);
    }
    printf("summary  %zu motions, %zu reparents, %zu states checked, %zu colliding\n", cx.n_motion, cx.n_reparent, cx.n_states, cx.n_bad);
    tmplan_destroy(plan);
    tmplan_destroy(start);
    aa_rx_sg_destroy(cx.sg);
    free(cx.q);
    return cx.n_bad ? 1 : 0;
}
