/* planfile.c - plan files: the on-disk format either side of the motion-planning path.
 *
 * Format (reference: doc/md/planfile.md:13-76; scanner states src/planfile.l:77-192; writer
 * src/tmplan.c:253-296 and lisp/tm-plan.lisp:119-152): one statement per line,
 *     a ACTION arg ...        task action
 *     r frame new_parent      reparent
 *     m name ...              motion plan start: configuration variable names
 *     p value ...             waypoint (one value per name of the preceding m line)
 *     # comment
 * This is an independent implementation with its own in-memory layout (the reference's lives in
 * Amino memory regions); only the text format is shared, so files written here replay in the
 * reference's tmsmt --gui / tmpexec and the reference's files (e.g. demo/baxter-sussman/q0.tmp)
 * parse here.
 */
#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tmk_internal.h"

typedef struct {
    int kind;       /* 'a', 'r', 'm' */
    size_t n_text;  /* a: 1 | r: 2 | m: number of names */
    char **text;
    size_t n_points;
    double *points; /* m: n_points x n_text */
    size_t n_flat;  /* values added one by one through tmplan_op_motion_plan_path_add */
} plan_op;

struct tmk_plan {
    size_t n, cap;
    plan_op **ops; /* ops are allocated one by one: their addresses are handed out by the tmplan_* API */
};

static char *dup_n(const char *s, size_t n) {
    char *d = (char *)malloc(n + 1);
    memcpy(d, s, n);
    d[n] = 0;
    return d;
}

static plan_op *push_op(struct tmk_plan *p, int kind) {
    if (p->n == p->cap) {
        p->cap = p->cap ? 2 * p->cap : 16;
        p->ops = (plan_op **)realloc(p->ops, sizeof(plan_op *) * p->cap);
    }
    plan_op *o = (plan_op *)calloc(1, sizeof(plan_op));
    p->ops[p->n++] = o;
    o->kind = kind;
    return o;
}

struct tmk_plan *tmk_plan_create(void) { return (struct tmk_plan *)calloc(1, sizeof(struct tmk_plan)); }

void tmk_plan_destroy(struct tmk_plan *p) {
    if (!p) return;
    for (size_t i = 0; i < p->n; i++) {
        for (size_t k = 0; k < p->ops[i]->n_text; k++) free(p->ops[i]->text[k]);
        free(p->ops[i]->text);
        free(p->ops[i]->points);
        free(p->ops[i]);
    }
    free(p->ops);
    free(p);
}

void tmk_plan_add_action(struct tmk_plan *p, const char *text) {
    plan_op *o = push_op(p, 'a');
    o->n_text = 1;
    o->text = (char **)malloc(sizeof(char *));
    while (*text && isspace((unsigned char)*text)) text++;
    size_t n = strlen(text);
    while (n && isspace((unsigned char)text[n - 1])) n--;
    o->text[0] = dup_n(text, n);
}

void tmk_plan_add_reparent(struct tmk_plan *p, const char *frame, const char *new_parent) {
    plan_op *o = push_op(p, 'r');
    o->n_text = 2;
    o->text = (char **)malloc(2 * sizeof(char *));
    o->text[0] = dup_n(frame, strlen(frame));
    o->text[1] = dup_n(new_parent, strlen(new_parent));
}

void tmk_plan_add_motion(struct tmk_plan *p, size_t n_names, const char *const *names, size_t n_points, const double *points,
                         size_t ld) {
    if (ld < n_names) tmk_die("tmk_plan_add_motion: ld < number of names");
    plan_op *o = push_op(p, 'm');
    o->n_text = n_names;
    o->text = (char **)malloc(sizeof(char *) * (n_names + 1));
    for (size_t k = 0; k < n_names; k++) o->text[k] = dup_n(names[k], strlen(names[k]));
    o->n_points = n_points;
    o->points = (double *)malloc(sizeof(double) * (n_points * n_names + 1));
    for (size_t i = 0; i < n_points; i++) memcpy(o->points + i * n_names, points + i * ld, sizeof(double) * n_names);
    o->n_flat = n_points * n_names;
}

static void write_ops(const struct tmk_plan *p, FILE *f) {
    for (size_t i = 0; i < p->n; i++) {
        const plan_op *o = p->ops[i];
        if (o->kind == 'a') {
            if (o->n_text && o->text[0][0]) fprintf(f, "a %s\n", o->text[0]);
            else fprintf(f, "a\n");
        }
        else if (o->kind == 'r') fprintf(f, "r %s %s\n", o->n_text > 0 ? o->text[0] : "", o->n_text > 1 ? o->text[1] : "");
        else {
            fputc('m', f);
            for (size_t k = 0; k < o->n_text; k++) fprintf(f, " %s", o->text[k]);
            fputc('\n', f);
            for (size_t r = 0; r < o->n_points; r++) {
                fputc('p', f);
                for (size_t k = 0; k < o->n_text; k++) fprintf(f, " %.17g", o->points[r * o->n_text + k]);
                fputc('\n', f);
            }
        }
    }
}

int tmk_plan_write(const struct tmk_plan *p, const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) {
        fprintf(stderr, "tmkcl: could not open plan file `%s' for writing: %s\n", path, strerror(errno));
        return -1;
    }
    write_ops(p, f);
    if (fclose(f)) return -1;
    return 0;
}

/* next whitespace-delimited token of *s; returns its length, 0 at the end of the line / at a comment */
static size_t next_token(const char **s, const char **tok) {
    const char *c = *s;
    while (*c == ' ' || *c == '\t' || *c == '\r') c++;
    if (!*c || *c == '\n' || *c == '#') { *s = c; return 0; }
    *tok = c;
    while (*c && !isspace((unsigned char)*c) && *c != '#') c++;
    *s = c;
    return (size_t)(c - *tok);
}

static struct tmk_plan *parse_stream(FILE *f, const char *path);

struct tmk_plan *tmk_plan_parse(const char *path) {
    FILE *f = fopen(path, "r");
    if (!f) {
        fprintf(stderr, "tmkcl: could not open plan file `%s': %s\n", path, strerror(errno));
        return NULL;
    }
    struct tmk_plan *p = parse_stream(f, path);
    fclose(f);
    return p;
}

static struct tmk_plan *parse_stream(FILE *f, const char *path) {
    struct tmk_plan *p = tmk_plan_create();
    char *line = NULL;
    size_t cap = 0, lineno = 0;
    ssize_t len;
    int bad = 0;
    while (!bad && (len = getline(&line, &cap, f)) >= 0) {
        lineno++;
        const char *s = line, *tok;
        size_t n = next_token(&s, &tok);
        if (!n) continue; /* blank or comment */
        if (n != 1) { bad = 1; break; }
        char kind = tok[0];
        if (kind == 'a') {
            const char *rest = s;
            const char *hash = strchr(rest, '#');
            size_t m = hash ? (size_t)(hash - rest) : strlen(rest);
            char *t = dup_n(rest, m);
            tmk_plan_add_action(p, t);
            free(t);
        } else if (kind == 'r') {
            const char *a, *b;
            size_t na = next_token(&s, &a), nb = na ? next_token(&s, &b) : 0;
            if (!na || !nb) { bad = 1; break; }
            char *fa = dup_n(a, na), *fb = dup_n(b, nb);
            tmk_plan_add_reparent(p, fa, fb);
            free(fa); free(fb);
        } else if (kind == 'm') {
            plan_op *o = push_op(p, 'm');
            size_t capn = 8;
            o->text = (char **)malloc(sizeof(char *) * capn);
            while ((n = next_token(&s, &tok))) {
                if (o->n_text == capn) { capn *= 2; o->text = (char **)realloc(o->text, sizeof(char *) * capn); }
                o->text[o->n_text++] = dup_n(tok, n);
            }
            if (!o->n_text) bad = 1;
        } else if (kind == 'p') {
            if (!p->n || p->ops[p->n - 1]->kind != 'm') { bad = 1; break; } /* waypoint outside a motion plan */
            plan_op *o = p->ops[p->n - 1];
            o->points = (double *)realloc(o->points, sizeof(double) * (o->n_points + 1) * o->n_text);
            size_t k = 0;
            while ((n = next_token(&s, &tok))) {
                char *end = NULL;
                double v = strtod(tok, &end);
                if (end != tok + n || k >= o->n_text) { bad = 1; break; }
                o->points[o->n_points * o->n_text + k++] = v;
            }
            if (k != o->n_text) bad = 1; /* every waypoint has as many values as names (lisp/tm-plan.lisp:132-136) */
            o->n_points++;
            o->n_flat = o->n_points * o->n_text;
        } else bad = 1;
    }
    free(line);
    if (bad) {
        fprintf(stderr, "tmkcl: plan file `%s': syntax error on line %zu\n", path, lineno);
        tmk_plan_destroy(p);
        return NULL;
    }
    return p;
}

size_t tmk_plan_op_count(const struct tmk_plan *p) { return p->n; }
int tmk_plan_op_kind(const struct tmk_plan *p, size_t i) {
    if (i >= p->n) tmk_die("tmk_plan_op_kind: index out of range");
    return p->ops[i]->kind;
}
const char *tmk_plan_op_text(const struct tmk_plan *p, size_t i, size_t k) {
    if (i >= p->n || k >= p->ops[i]->n_text) tmk_die("tmk_plan_op_text: index out of range");
    return p->ops[i]->text[k];
}
size_t tmk_plan_motion_names(const struct tmk_plan *p, size_t i) {
    if (i >= p->n || p->ops[i]->kind != 'm') tmk_die("tmk_plan_motion_names: not a motion plan");
    return p->ops[i]->n_text;
}
size_t tmk_plan_motion_points(const struct tmk_plan *p, size_t i) {
    if (i >= p->n || p->ops[i]->kind != 'm') tmk_die("tmk_plan_motion_points: not a motion plan");
    return p->ops[i]->n_points;
}
const double *tmk_plan_motion_data(const struct tmk_plan *p, size_t i) {
    if (i >= p->n || p->ops[i]->kind != 'm') tmk_die("tmk_plan_motion_data: not a motion plan");
    return p->ops[i]->points;
}

/* ------------------------------------------------------------------------------------------------
 * The reference's own names for the same thing: the tmplan_* C ABI of libtmsmt
 * (/root/reference/include/tmsmt/tmplan.h:34-247), called by demo/tmpexec.c:111-125, 246-312 and bound from
 * Lisp at lisp/foreign-tmplan.lisp:16-50.  Thin shims over the structures above: struct tmplan IS struct
 * tmk_plan, struct tmplan_op IS one plan_op.  The reference allocates out of an Amino memory region
 * (struct aa_mem_region, not in the reference tree): here the region argument is accepted and ignored (it may
 * be NULL) and the plan lives on the heap until tmplan_destroy (added).
 * ---------------------------------------------------------------------------------------------- */
#include "tmkit_b200/tmplan.h"

struct tmplan *tmplan_create(struct aa_mem_region *reg) { (void)reg; return (struct tmplan *)tmk_plan_create(); }
void tmplan_destroy(struct tmplan *tmp) { tmk_plan_destroy((struct tmk_plan *)tmp); }
size_t tmplan_op_count(struct tmplan *tmp) { return ((struct tmk_plan *)tmp)->n; }
void tmplan_finish_op(struct tmplan *tmp) { (void)tmp; } /* ops are complete as soon as they are added */

void tmplan_add_op(struct tmplan *tmp, void *op) {
    struct tmk_plan *p = (struct tmk_plan *)tmp;
    if (p->n == p->cap) {
        p->cap = p->cap ? 2 * p->cap : 16;
        p->ops = (plan_op **)realloc(p->ops, sizeof(plan_op *) * p->cap);
    }
    p->ops[p->n++] = (plan_op *)op; /* ownership passes to the plan */
}
void tmplan_add_action(struct tmplan *tmp) { push_op((struct tmk_plan *)tmp, 'a'); }
void tmplan_add_motion_plan(struct tmplan *tmp) { push_op((struct tmk_plan *)tmp, 'm'); }
void tmplan_add_reparent(struct tmplan *tmp) {
    plan_op *o = push_op((struct tmk_plan *)tmp, 'r');
    o->text = (char **)calloc(2, sizeof(char *));
    o->text[0] = dup_n("", 0);
    o->text[1] = dup_n("", 0);
    o->n_text = 2;
}
struct tmplan_op *tmplan_last_op(struct tmplan *tmp) {
    struct tmk_plan *p = (struct tmk_plan *)tmp;
    return p->n ? (struct tmplan_op *)p->ops[p->n - 1] : NULL;
}
enum tmplan_op_type tmplan_op_type(const struct tmplan_op *op) {
    switch (((const plan_op *)op)->kind) {
    case 'a': return TMPLAN_OP_ACTION;
    case 'm': return TMPLAN_OP_MOTION_PLAN;
    default: return TMPLAN_OP_REPARENT;
    }
}

static void set_text(plan_op *o, size_t k, const char *value) {
    if (k >= o->n_text) {
        o->text = (char **)realloc(o->text, sizeof(char *) * (k + 1));
        for (size_t j = o->n_text; j <= k; j++) o->text[j] = dup_n("", 0);
        o->n_text = k + 1;
    }
    free(o->text[k]);
    o->text[k] = dup_n(value ? value : "", value ? strlen(value) : 0);
}
void tmplan_op_action_set(struct tmplan_op_action *op, const char *value) { set_text((plan_op *)op, 0, value); }
const char *tmplan_op_action_get(const struct tmplan_op_action *op) {
    const plan_op *o = (const plan_op *)op;
    return o->n_text ? o->text[0] : NULL;
}

void tmplan_op_motion_plan_add_var(struct tmplan_op_motion_plan *op, const char *name) {
    plan_op *o = (plan_op *)op;
    if (o->n_points) tmk_die("tmplan_op_motion_plan_add_var: the path has points already");
    set_text(o, o->n_text, name);
}
void tmplan_op_motion_plan_map_var(struct tmplan_op_motion_plan *op, void (*function)(void *cx, const char *name), void *cx) {
    const plan_op *o = (const plan_op *)op;
    for (size_t k = 0; k < o->n_text; k++) function(cx, o->text[k]);
}
void tmplan_op_motion_plan_vars(struct tmplan_op_motion_plan *op, size_t n, const char **name) {
    const plan_op *o = (const plan_op *)op;
    for (size_t k = 0; k < n && k < o->n_text; k++) name[k] = o->text[k];
}
int tmplan_map_ops(const struct tmplan *tmp, void (*function)(void *cx, const struct tmplan_op *op), void *cx) {
    const struct tmk_plan *p = (const struct tmk_plan *)tmp;
    for (size_t i = 0; i < p->n; i++) function(cx, (const struct tmplan_op *)p->ops[i]);
    return 0;
}

/* the path is a flat array of values; its length need not be a multiple of the variable count while it is
 * being built (the reference's scanner adds value by value, src/planfile.l:150-170) */
void tmplan_op_motion_plan_path_start(struct tmplan_op_motion_plan *op) { (void)op; }
void tmplan_op_motion_plan_path_add(struct tmplan_op_motion_plan *op, double x) {
    plan_op *o = (plan_op *)op;
    if (!o->n_text) tmk_die("tmplan_op_motion_plan_path_add: no configuration variables");
    o->points = (double *)realloc(o->points, sizeof(double) * (o->n_flat + 1));
    o->points[o->n_flat++] = x;
    o->n_points = o->n_flat / o->n_text;
}
void tmplan_op_motion_plan_path_finish(struct tmplan_op_motion_plan *op) {
    plan_op *o = (plan_op *)op;
    if (o->n_text && o->n_flat % o->n_text) tmk_die("tmplan: path size %zu is not a multiple of %zu variables", o->n_flat, o->n_text);
}
double *tmplan_op_motion_plan_path(struct tmplan_op_motion_plan *op) { return ((plan_op *)op)->points; }
size_t tmplan_op_motion_plan_config_count(struct tmplan_op_motion_plan *op) { return ((plan_op *)op)->n_text; }
size_t tmplan_op_motion_plan_path_size(struct tmplan_op_motion_plan *op) {
    const plan_op *o = (const plan_op *)op;
    return o->n_points * o->n_text;
}

void tmplan_op_reparent_set_frame(struct tmplan_op_reparent *op, const char *frame) { set_text((plan_op *)op, 0, frame); }
void tmplan_op_reparent_set_new_parent(struct tmplan_op_reparent *op, const char *frame) { set_text((plan_op *)op, 1, frame); }
const char *tmplan_op_reparent_get_frame(struct tmplan_op_reparent *op) { return ((plan_op *)op)->text[0]; }
const char *tmplan_op_reparent_get_new_parent(struct tmplan_op_reparent *op) { return ((plan_op *)op)->text[1]; }

struct tmplan *tmplan_parse_file(FILE *in, struct aa_mem_region *reg) {
    (void)reg;
    return (struct tmplan *)parse_stream(in, "<stream>");
}
struct tmplan *tmplan_parse_filename(const char *filename, struct aa_mem_region *reg) {
    (void)reg;
    FILE *in = fopen(filename, "r");
    if (!in) {
        fprintf(stderr, "Could not open `%s'.", filename); /* the reference's message, src/tmplan.c:301-305 */
        return NULL;
    }
    struct tmplan *r = tmplan_parse_file(in, reg);
    fclose(in);
    return r;
}
int tmplan_write(const struct tmplan *tmp, FILE *out) {
    write_ops((const struct tmk_plan *)tmp, out);
    return ferror(out) ? -1 : 0;
}
