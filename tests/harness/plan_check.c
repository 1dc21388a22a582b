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

static void names_to_ids(const struct aa_rx_sg *sg, const struct tmk_plan *p, size_t op, size_t n, aa_rx_config_id *ids) {
    const char **names = (const char **)malloc(sizeof(char *) * n);
    for (size_t k = 0; k < n; k++) names[k] = tmk_plan_op_text(p, op, k);
    aa_rx_sg_config_indices(sg, n, names, ids); /* map by NAME: tmpexec.c:269-272 */
    free(names);
}

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

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s plugin.so scene start.tmp plan.tmp [resolution]\n", argv[0]);
        return 2;
    }
    double resolution = argc > 5 ? atof(argv[5]) : 0.01; /* OMPL longestValidSegmentFraction */
    struct aa_rx_sg *sg = aa_rx_dl_sg(argv[1], argv[2], NULL);
    if (!sg) return 2;
    aa_rx_sg_init(sg);
    aa_rx_sg_cl_init(sg);
    size_t n_q = aa_rx_sg_config_count(sg);
    double *q = (double *)calloc(n_q, sizeof(double));

    /* start configuration from a plan file with one motion plan of one waypoint (-q q0.tmp) */
    struct tmk_plan *start = tmk_plan_parse(argv[3]);
    if (!start || tmk_plan_op_count(start) < 1 || tmk_plan_op_kind(start, 0) != 'm') return 2;
    {
        size_t n = tmk_plan_motion_names(start, 0);
        aa_rx_config_id *ids = (aa_rx_config_id *)malloc(sizeof(int) * n);
        names_to_ids(sg, start, 0, n, ids);
        size_t last = tmk_plan_motion_points(start, 0) - 1;
        aa_rx_sg_config_set(sg, n_q, n, ids, tmk_plan_motion_data(start, 0) + last * n, q);
        free(ids);
    }
    struct tmk_plan *plan = tmk_plan_parse(argv[4]);
    if (!plan) return 2;

    size_t n_states = 0, n_bad = 0, n_motion = 0, n_reparent = 0;
    for (size_t op = 0; op < tmk_plan_op_count(plan); op++) {
        int kind = tmk_plan_op_kind(plan, op);
        if (kind == 'a') {
            printf("action   %s\n", tmk_plan_op_text(plan, op, 0));
        } else if (kind == 'm') {
            size_t n = tmk_plan_motion_names(plan, op), np = tmk_plan_motion_points(plan, op);
            const double *pts = tmk_plan_motion_data(plan, op);
            aa_rx_config_id *ids = (aa_rx_config_id *)malloc(sizeof(int) * n);
            names_to_ids(sg, plan, op, n, ids);
            /* the scene allows what collides at the start of the motion (tmpexec.c:118) */
            aa_rx_sg_allow_config(sg, n_q, q);
            struct aa_rx_cl *cl = aa_rx_cl_create(sg);
            struct aa_rx_cl_set *cl_set = aa_rx_cl_set_create(sg);
            /* every waypoint with the reference's single-state check ... */
            size_t bad = 0;
            double *qq = (double *)malloc(sizeof(double) * n_q);
            memcpy(qq, q, sizeof(double) * n_q);
            for (size_t i = 0; i < np; i++) {
                aa_rx_sg_config_set(sg, n_q, n, ids, pts + i * n, qq);
                bad += check_state(sg, cl, cl_set, qq, 1) != 0;
                n_states++;
            }
            /* ... and every segment at the planner's resolution through the swept-edge entry point */
            if (np > 1 && resolution > 0) {
                double ext2 = 0;
                for (size_t c = 0; c < n; c++) {
                    double lo = -M_PI, hi = M_PI;
                    aa_rx_sg_get_limit_pos(sg, ids[c], &lo, &hi);
                    ext2 += (hi - lo) * (hi - lo);
                }
                uint32_t *bits = (uint32_t *)calloc((np + 31) / 32 + 1, sizeof(uint32_t));
                bad += aa_rx_cl_check_edges(cl, np - 1, n, ids, n_q, q, pts, pts + n, n, resolution * sqrt(ext2), bits, NULL);
                n_states += np - 1;
                free(bits);
            }
            printf("motion   %zu waypoints over %zu joints: %zu colliding states\n", np, n, bad);
            n_bad += bad;
            n_motion++;
            if (np) aa_rx_sg_config_set(sg, n_q, n, ids, pts + (np - 1) * n, q);
            free(qq); free(ids);
            aa_rx_cl_set_destroy(cl_set);
            aa_rx_cl_destroy(cl);
        } else if (kind == 'r') {
            /* reparent at the current configuration, world pose preserved (tmpexec.c:307-348) */
            const char *frame = tmk_plan_op_text(plan, op, 0), *parent = tmk_plan_op_text(plan, op, 1);
            size_t n_f = aa_rx_sg_frame_count(sg);
            double *TF_abs = (double *)malloc(sizeof(double) * 7 * n_f), rel[7];
            aa_rx_sg_tf(sg, n_q, q, n_f, NULL, 7, TF_abs, 7);
            aa_rx_sg_rel_tf(sg, aa_rx_sg_frame_id(sg, parent), aa_rx_sg_frame_id(sg, frame), TF_abs, 7, rel);
            struct aa_rx_sg *sg2 = aa_rx_sg_copy(sg);
            aa_rx_sg_reparent_name(sg2, parent, frame, rel);
            aa_rx_sg_init(sg2);
            aa_rx_sg_cl_init(sg2);
            aa_rx_sg_destroy(sg);
            sg = sg2;
            free(TF_abs);
            printf("reparent %s -> %s\n", frame, parent);
            n_reparent++;
        }
    }
    /* final world poses of the blocks, for the caller to compare */
    {
        size_t n_f = aa_rx_sg_frame_count(sg);
        double *TF_abs = (double *)malloc(sizeof(double) * 7 * n_f);
        aa_rx_sg_tf(sg, n_q, q, n_f, NULL, 7, TF_abs, 7);
        for (size_t i = 0; i < n_f; i++)
            if (0 == strncmp(aa_rx_sg_frame_name(sg, (int)i), "block_", 6))
                printf("final    %s %.9f %.9f %.9f\n", aa_rx_sg_frame_name(sg, (int)i), TF_abs[7 * i + 4], TF_abs[7 * i + 5], TF_abs[7 * i + 6]);
        free(TF_abs);
    }
    printf("summary  %zu motions, %zu reparents, %zu states checked, %zu colliding\n", n_motion, n_reparent, n_states, n_bad);
    tmk_plan_destroy(plan);
    tmk_plan_destroy(start);
    aa_rx_sg_destroy(sg);
    free(q);
    return n_bad ? 1 : 0;
}
