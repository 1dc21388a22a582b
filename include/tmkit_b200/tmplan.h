/* tmplan.h - the reference's plan-file C ABI (libtmsmt), exported by libtmkcl.so under the reference's names.
 *
 * Declarations follow /root/reference/include/tmsmt/tmplan.h:34-247 one to one (same names, argument lists and
 * meaning); callers in the reference: demo/tmpexec.c:111-125 (parse, count, map), :246-312 (op type, action text,
 * motion-plan path / variables, reparent frames), lisp/foreign-tmplan.lisp:16-50 (CFFI defcfuns), :68-72,123-153
 * (callbacks through tmplan_map_ops / tmplan_op_motion_plan_map_var).
 *
 * Difference, stated: the reference allocates a plan out of an Amino memory region (struct aa_mem_region; Amino is
 * not in the reference tree) and frees it by popping the region.  Here the region argument is accepted and ignored
 * (NULL is fine), the plan lives on the heap, and tmplan_destroy (added) frees it.  The writer prints waypoints
 * with %.17g (exact round trip) where the reference prints %f (src/tmplan.c:281).
 */
#ifndef TMKIT_B200_TMPLAN_H
#define TMKIT_B200_TMPLAN_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

enum tmplan_op_type {       /* tmplan.h:13-17 */
    TMPLAN_OP_ACTION,       /* a task action */
    TMPLAN_OP_MOTION_PLAN,  /* a motion plan */
    TMPLAN_OP_REPARENT      /* a scene-graph reparenting */
};

struct aa_mem_region; /* opaque, ignored */
struct tmplan;
struct tmplan_op;
struct tmplan_op_action;
struct tmplan_op_motion_plan;
struct tmplan_op_reparent;

struct tmplan *tmplan_create(struct aa_mem_region *reg);                        /* tmplan.h:34-35 */
void tmplan_destroy(struct tmplan *tmp);                                        /* added (no region to pop) */
size_t tmplan_op_count(struct tmplan *tmp);                                     /* tmplan.h:41-42; tmpexec.c:121 */
void tmplan_finish_op(struct tmplan *tmp);                                      /* tmplan.h:49-50 */
void tmplan_add_op(struct tmplan *tmp, void *op);                               /* tmplan.h:60-61 */
void tmplan_add_action(struct tmplan *tmp);                                     /* tmplan.h:66-67 */
void tmplan_add_motion_plan(struct tmplan *tmp);                                /* tmplan.h:72-73 */
void tmplan_add_reparent(struct tmplan *tmp);                                   /* tmplan.h:78-79 */
struct tmplan_op *tmplan_last_op(struct tmplan *tmp);                           /* tmplan.h:94-95 */
enum tmplan_op_type tmplan_op_type(const struct tmplan_op *op);                 /* tmplan.h:100-101; tmpexec.c:246 */

void tmplan_op_action_set(struct tmplan_op_action *op, const char *value);      /* tmplan.h:113-114 */
const char *tmplan_op_action_get(const struct tmplan_op_action *op);            /* tmplan.h:119-120; tmpexec.c:254 */

void tmplan_op_motion_plan_add_var(struct tmplan_op_motion_plan *op, const char *name);          /* tmplan.h:132-133 */
void tmplan_op_motion_plan_map_var(struct tmplan_op_motion_plan *op, void (*function)(void *cx, const char *name),
                                   void *cx);                                                    /* tmplan.h:138-141 */
void tmplan_op_motion_plan_vars(struct tmplan_op_motion_plan *op, size_t n, const char **name);  /* :146-148; tmpexec.c:271 */
int tmplan_map_ops(const struct tmplan *tmp, void (*function)(void *cx, const struct tmplan_op *op), void *cx);
                                                                                                 /* :153-156; tmpexec.c:125 */
void tmplan_op_motion_plan_path_start(struct tmplan_op_motion_plan *op);        /* tmplan.h:161-162 */
void tmplan_op_motion_plan_path_add(struct tmplan_op_motion_plan *op, double x); /* tmplan.h:167-168 */
void tmplan_op_motion_plan_path_finish(struct tmplan_op_motion_plan *op);       /* tmplan.h:173-174 */
double *tmplan_op_motion_plan_path(struct tmplan_op_motion_plan *op);           /* tmplan.h:179-180; tmpexec.c:265 */
size_t tmplan_op_motion_plan_config_count(struct tmplan_op_motion_plan *op);    /* tmplan.h:186-187; tmpexec.c:266 */
size_t tmplan_op_motion_plan_path_size(struct tmplan_op_motion_plan *op);       /* tmplan.h:196-197; tmpexec.c:267 */

void tmplan_op_reparent_set_frame(struct tmplan_op_reparent *op, const char *frame);       /* tmplan.h:209-210 */
void tmplan_op_reparent_set_new_parent(struct tmplan_op_reparent *op, const char *frame);  /* tmplan.h:215-216 */
const char *tmplan_op_reparent_get_frame(struct tmplan_op_reparent *op);                   /* :221-222; tmpexec.c:310 */
const char *tmplan_op_reparent_get_new_parent(struct tmplan_op_reparent *op);              /* :227-228; tmpexec.c:311 */

struct tmplan *tmplan_parse_filename(const char *filename, struct aa_mem_region *reg);  /* :233-234; tmpexec.c:111 */
struct tmplan *tmplan_parse_file(FILE *in, struct aa_mem_region *reg);                  /* tmplan.h:239-240 */
int tmplan_write(const struct tmplan *tmp, FILE *out);                                  /* tmplan.h:246-247 */

#ifdef __cplusplus
}
#endif
#endif
