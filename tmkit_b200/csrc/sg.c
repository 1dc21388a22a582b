/* sg.c - host side of the scene graph: construction, sorting, naming, mutation.
 *
 * Mirrors the Amino scene-graph interface at the call shapes the reference uses
 * (demo/tmpexec.c:94-150, 264-348, 362-428; lisp/tm-plan.lisp:50-59).  Nothing here is
 * per-configuration work: FK (aa_rx_sg_tf) and collision run on the GPU (cl.cu).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tmk_internal.h"

void tmk_die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fputs("tmkcl: ", stderr);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
    abort();
}

static char *dupstr(const char *s) {
    if (!s) s = "";
    size_t n = strlen(s) + 1;
    char *d = (char *)malloc(n);
    memcpy(d, s, n);
    return d;
}

#define GROW(ptr, n, cap, type)                                                  \
    do {                                                                         \
        if ((n) == (cap)) {                                                      \
            (cap) = (cap) ? 2 * (cap) : 16;                                      \
            (ptr) = (type *)realloc((ptr), (cap) * sizeof(type));                \
        }                                                                        \
    } while (0)

/* ------------------------------------------------------------------ transforms (double) */
void tmk_qmul(const double *a, const double *b, double *c) {
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    c[0] = x; c[1] = y; c[2] = z; c[3] = w;
}
void tmk_qrot(const double *q, const double *v, double *out) {
    double tx = 2 * (q[1] * v[2] - q[2] * v[1]);
    double ty = 2 * (q[2] * v[0] - q[0] * v[2]);
    double tz = 2 * (q[0] * v[1] - q[1] * v[0]);
    double x = v[0] + q[3] * tx + (q[1] * tz - q[2] * ty);
    double y = v[1] + q[3] * ty + (q[2] * tx - q[0] * tz);
    double z = v[2] + q[3] * tz + (q[0] * ty - q[1] * tx);
    out[0] = x; out[1] = y; out[2] = z;
}
void tmk_tfmul(const double *a, const double *b, double *c) {
    double r[4], v[3];
    tmk_qmul(a, b, r);
    tmk_qrot(a, b + 4, v);
    c[0] = r[0]; c[1] = r[1]; c[2] = r[2]; c[3] = r[3];
    c[4] = v[0] + a[4]; c[5] = v[1] + a[5]; c[6] = v[2] + a[6];
}
void tmk_tfinv(const double *a, double *c) {
    double qi[4] = {-a[0], -a[1], -a[2], a[3]}, v[3];
    tmk_qrot(qi, a + 4, v);
    c[0] = qi[0]; c[1] = qi[1]; c[2] = qi[2]; c[3] = qi[3];
    c[4] = -v[0]; c[5] = -v[1]; c[6] = -v[2];
}

/* ------------------------------------------------------------------ geometry objects */
struct aa_rx_geom *aa_rx_geom_box(struct aa_rx_geom_opt *opt, const double dimension[3]) {
    (void)opt;
    struct aa_rx_geom *g = (struct aa_rx_geom *)calloc(1, sizeof(*g));
    g->shape = TMK_BOX;
    memcpy(g->dim, dimension, sizeof(double) * 3);
    if (!(g->dim[0] > 0 && g->dim[1] > 0 && g->dim[2] > 0)) tmk_die("aa_rx_geom_box: non-positive dimension");
    return g;
}
struct aa_rx_geom *aa_rx_geom_sphere(struct aa_rx_geom_opt *opt, double radius) {
    (void)opt;
    struct aa_rx_geom *g = (struct aa_rx_geom *)calloc(1, sizeof(*g));
    g->shape = TMK_SPHERE;
    g->dim[0] = radius;
    if (!(radius > 0)) tmk_die("aa_rx_geom_sphere: non-positive radius");
    return g;
}
struct aa_rx_geom *aa_rx_geom_cylinder(struct aa_rx_geom_opt *opt, double height, double radius) {
    (void)opt;
    struct aa_rx_geom *g = (struct aa_rx_geom *)calloc(1, sizeof(*g));
    g->shape = TMK_CYLINDER;
    g->dim[0] = height;
    g->dim[1] = radius;
    if (!(radius > 0 && height > 0)) tmk_die("aa_rx_geom_cylinder: non-positive size");
    return g;
}
struct aa_rx_mesh *aa_rx_mesh_create(void) {
    struct aa_rx_mesh *m = (struct aa_rx_mesh *)calloc(1, sizeof(*m));
    m->refcount = 1;
    return m;
}
void aa_rx_mesh_destroy(struct aa_rx_mesh *m) {
    if (!m) return;
    if (--m->refcount > 0) return;
    tmk_mesh_compiled_release(m->compiled);
    free(m->verts);
    free(m->idx);
    free(m);
}
void aa_rx_mesh_set_vertices(struct aa_rx_mesh *m, size_t n, const float *v, int copy) {
    (void)copy; /* always copied: the handle owns its data */
    tmk_mesh_compiled_release(m->compiled);
    m->compiled = NULL;
    free(m->verts);
    m->nv = n;
    m->verts = (float *)malloc(sizeof(float) * 3 * n);
    memcpy(m->verts, v, sizeof(float) * 3 * n);
}
void aa_rx_mesh_set_indices(struct aa_rx_mesh *m, size_t n, const unsigned *idx, int copy) {
    (void)copy;
    tmk_mesh_compiled_release(m->compiled);
    m->compiled = NULL;
    free(m->idx);
    m->nt = n;
    m->idx = (unsigned *)malloc(sizeof(unsigned) * 3 * n);
    memcpy(m->idx, idx, sizeof(unsigned) * 3 * n);
}
struct aa_rx_geom *aa_rx_geom_mesh(struct aa_rx_geom_opt *opt, struct aa_rx_mesh *mesh) {
    (void)opt;
    if (!mesh || !mesh->nv || !mesh->nt) tmk_die("aa_rx_geom_mesh: empty mesh");
    for (size_t i = 0; i < 3 * mesh->nt; i++)
        if (mesh->idx[i] >= mesh->nv) tmk_die("aa_rx_geom_mesh: index %u out of range", mesh->idx[i]);
    struct aa_rx_geom *g = (struct aa_rx_geom *)calloc(1, sizeof(*g));
    g->shape = TMK_MESH;
    g->mesh = mesh;
    mesh->refcount++;
    return g;
}
static void geom_free(struct aa_rx_geom *g) {
    if (g->mesh) aa_rx_mesh_destroy(g->mesh);
    free(g);
}
static struct aa_rx_geom *geom_clone(const struct aa_rx_geom *g) {
    struct aa_rx_geom *c = (struct aa_rx_geom *)malloc(sizeof(*c));
    *c = *g;
    if (c->mesh) c->mesh->refcount++;
    return c;
}

/* ------------------------------------------------------------------ construction */
struct aa_rx_sg *aa_rx_sg_create(void) { return (struct aa_rx_sg *)calloc(1, sizeof(struct aa_rx_sg)); }

static void compiled_free(struct aa_rx_sg *sg) {
    free(sg->order); free(sg->parent); free(sg->type); free(sg->E); free(sg->axis); free(sg->offset);
    free(sg->qidx); free(sg->config_names); free(sg->lim_min); free(sg->lim_max); free(sg->allowed);
    free(sg->g_frame); free(sg->g_shape); free(sg->g_mesh); free(sg->g_dim); free(sg->meshes);
    sg->order = sg->parent = sg->type = sg->qidx = sg->g_frame = sg->g_shape = sg->g_mesh = NULL;
    sg->E = sg->axis = sg->offset = sg->lim_min = sg->lim_max = sg->g_dim = NULL;
    sg->config_names = NULL; sg->allowed = NULL; sg->meshes = NULL;
    if (sg->dev) { tmk_dev_sg_release(sg->dev); sg->dev = NULL; }
    sg->inited = sg->cl_inited = 0;
}

void aa_rx_sg_destroy(struct aa_rx_sg *sg) {
    if (!sg) return;
    compiled_free(sg);
    for (size_t i = 0; i < sg->n_def; i++) {
        tmk_frame_def *d = &sg->defs[i];
        free(d->name); free(d->parent); free(d->config);
        for (int k = 0; k < d->n_geom; k++) geom_free(d->geoms[k]);
        free(d->geoms);
    }
    free(sg->defs);
    for (size_t i = 0; i < sg->n_limit; i++) free(sg->limits[i].name);
    free(sg->limits);
    for (size_t i = 0; i < sg->n_allow; i++) { free(sg->allows[i].a); free(sg->allows[i].b); }
    free(sg->allows);
    free(sg);
}

static tmk_frame_def *find_def(const struct aa_rx_sg *sg, const char *name) {
    for (size_t i = 0; i < sg->n_def; i++)
        if (0 == strcmp(sg->defs[i].name, name)) return &sg->defs[i];
    return NULL;
}

static tmk_frame_def *add_def(struct aa_rx_sg *sg, const char *parent, const char *name, const double q[4],
                              const double v[3]) {
    if (!name || !*name) tmk_die("frame needs a name");
    if (find_def(sg, name)) tmk_die("duplicate frame `%s'", name);
    GROW(sg->defs, sg->n_def, sg->cap_def, tmk_frame_def);
    tmk_frame_def *d = &sg->defs[sg->n_def++];
    memset(d, 0, sizeof(*d));
    d->name = dupstr(name);
    d->parent = dupstr(parent);
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (!(n > 0)) tmk_die("frame `%s': zero quaternion", name);
    for (int i = 0; i < 4; i++) d->E[i] = q[i] / n;
    for (int i = 0; i < 3; i++) d->E[4 + i] = v[i];
    d->axis[2] = 1;
    sg->inited = 0;
    return d;
}

void aa_rx_sg_add_frame_fixed(struct aa_rx_sg *sg, const char *parent, const char *name, const double q[4],
                              const double v[3]) {
    add_def(sg, parent, name, q, v)->type = TMK_FIXED;
}

static void add_joint(struct aa_rx_sg *sg, int type, const char *parent, const char *name, const double q[4],
                      const double v[3], const char *config_name, const double axis[3], double offset) {
    tmk_frame_def *d = add_def(sg, parent, name, q, v);
    d->type = type;
    d->config = dupstr(config_name && *config_name ? config_name : name);
    double n = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    if (!(n > 0)) tmk_die("frame `%s': zero axis", name);
    for (int i = 0; i < 3; i++) d->axis[i] = axis[i] / n;
    d->offset = offset;
}
void aa_rx_sg_add_frame_revolute(struct aa_rx_sg *sg, const char *parent, const char *name, const double q[4],
                                 const double v[3], const char *config_name, const double axis[3], double offset) {
    add_joint(sg, TMK_REVOLUTE, parent, name, q, v, config_name, axis, offset);
}
void aa_rx_sg_add_frame_prismatic(struct aa_rx_sg *sg, const char *parent, const char *name, const double q[4],
                                  const double v[3], const char *config_name, const double axis[3], double offset) {
    add_joint(sg, TMK_PRISMATIC, parent, name, q, v, config_name, axis, offset);
}

void aa_rx_sg_set_limit_pos(struct aa_rx_sg *sg, const char *config_name, double min, double max) {
    for (size_t i = 0; i < sg->n_limit; i++)
        if (0 == strcmp(sg->limits[i].name, config_name)) {
            sg->limits[i].min = min; sg->limits[i].max = max;
            sg->inited = 0;
            return;
        }
    GROW(sg->limits, sg->n_limit, sg->cap_limit, tmk_limit);
    sg->limits[sg->n_limit].name = dupstr(config_name);
    sg->limits[sg->n_limit].min = min;
    sg->limits[sg->n_limit].max = max;
    sg->n_limit++;
    sg->inited = 0;
}

void aa_rx_geom_attach(struct aa_rx_sg *sg, const char *frame, struct aa_rx_geom *geom) {
    tmk_frame_def *d = find_def(sg, frame);
    if (!d) tmk_die("aa_rx_geom_attach: no frame `%s'", frame);
    d->geoms = (struct aa_rx_geom **)realloc(d->geoms, sizeof(*d->geoms) * (size_t)(d->n_geom + 1));
    d->geoms[d->n_geom++] = geom;
    sg->inited = 0;
}

static void allow_name(struct aa_rx_sg *sg, const char *a, const char *b, int allowed) {
    GROW(sg->allows, sg->n_allow, sg->cap_allow, tmk_allow);
    sg->allows[sg->n_allow].a = dupstr(a);
    sg->allows[sg->n_allow].b = dupstr(b);
    sg->allows[sg->n_allow].allowed = allowed;
    sg->n_allow++;
}

void aa_rx_sg_allow_collision_name(struct aa_rx_sg *sg, const char *f0, const char *f1, int allowed) {
    allow_name(sg, f0, f1, allowed);
    if (sg->inited) {
        int i = aa_rx_sg_frame_id(sg, f0), j = aa_rx_sg_frame_id(sg, f1);
        if (i < 0 || j < 0) tmk_die("allow_collision: unknown frame `%s' or `%s'", f0, f1);
        tmk_bit_put(sg->allowed, sg->wpr, (size_t)i, (size_t)j, allowed);
        tmk_bit_put(sg->allowed, sg->wpr, (size_t)j, (size_t)i, allowed);
        sg->serial++;
    }
}
void aa_rx_sg_allow_collision(struct aa_rx_sg *sg, aa_rx_frame_id i, aa_rx_frame_id j, int allowed) {
    if (!sg->inited) tmk_die("aa_rx_sg_allow_collision: scene graph not initialised");
    if (i < 0 || j < 0 || (size_t)i >= sg->n_f || (size_t)j >= sg->n_f) tmk_die("allow_collision: bad frame id");
    aa_rx_sg_allow_collision_name(sg, aa_rx_sg_frame_name(sg, i), aa_rx_sg_frame_name(sg, j), allowed);
}

/* ------------------------------------------------------------------ init: sort + index */
void aa_rx_sg_init(struct aa_rx_sg *sg) {
    if (sg->inited) return;
    compiled_free(sg);
    size_t n = sg->n_def;
    sg->n_f = n;
    sg->order = (int *)malloc(sizeof(int) * (n + 1));
    int *placed = (int *)malloc(sizeof(int) * (n + 1)); /* def index -> sorted id, -1 */
    for (size_t i = 0; i < n; i++) placed[i] = -1;
    /* stable topological order: repeated passes over insertion order, a frame is emitted as soon as
     * its parent has been (the same rule as tmkit_b200/scenes.py flatten) */
    size_t done = 0;
    while (done < n) {
        size_t before = done;
        for (size_t i = 0; i < n; i++) {
            if (placed[i] >= 0) continue;
            tmk_frame_def *d = &sg->defs[i];
            int ok = (d->parent[0] == 0);
            if (!ok) {
                tmk_frame_def *p = find_def(sg, d->parent);
                if (!p) tmk_die("frame `%s': unknown parent `%s'", d->name, d->parent);
                ok = placed[p - sg->defs] >= 0;
            }
            if (ok) { placed[i] = (int)done; sg->order[done++] = (int)i; }
        }
        if (done == before) tmk_die("scene graph has a cycle");
    }
    sg->parent = (int *)malloc(sizeof(int) * (n + 1));
    sg->type = (int *)malloc(sizeof(int) * (n + 1));
    sg->E = (double *)malloc(sizeof(double) * 7 * (n + 1));
    sg->axis = (double *)malloc(sizeof(double) * 3 * (n + 1));
    sg->offset = (double *)malloc(sizeof(double) * (n + 1));
    sg->qidx = (int *)malloc(sizeof(int) * (n + 1));
    sg->config_names = (char **)malloc(sizeof(char *) * (n + 1));
    sg->n_q = 0;
    for (size_t id = 0; id < n; id++) {
        tmk_frame_def *d = &sg->defs[sg->order[id]];
        sg->parent[id] = d->parent[0] ? placed[find_def(sg, d->parent) - sg->defs] : AA_RX_FRAME_ROOT;
        sg->type[id] = d->type;
        memcpy(sg->E + 7 * id, d->E, sizeof(double) * 7);
        memcpy(sg->axis + 3 * id, d->axis, sizeof(double) * 3);
        sg->offset[id] = d->offset;
        sg->qidx[id] = -1;
        if (d->type != TMK_FIXED) {
            size_t k;
            for (k = 0; k < sg->n_q; k++)
                if (0 == strcmp(sg->config_names[k], d->config)) break;
            if (k == sg->n_q) sg->config_names[sg->n_q++] = d->config;
            sg->qidx[id] = (int)k;
        }
    }
    sg->lim_min = (double *)malloc(sizeof(double) * (sg->n_q + 1));
    sg->lim_max = (double *)malloc(sizeof(double) * (sg->n_q + 1));
    for (size_t k = 0; k < sg->n_q; k++) {
        sg->lim_min[k] = sg->lim_max[k] = NAN;
        for (size_t i = 0; i < sg->n_limit; i++)
            if (0 == strcmp(sg->limits[i].name, sg->config_names[k])) {
                sg->lim_min[k] = sg->limits[i].min;
                sg->lim_max[k] = sg->limits[i].max;
            }
    }
    sg->wpr = (n + 31) / 32;
    sg->allowed = (uint32_t *)calloc(n * sg->wpr + 1, sizeof(uint32_t));
    sg->inited = 1; /* frame_id lookups below need it */
    for (size_t i = 0; i < sg->n_allow; i++) {
        int a = aa_rx_sg_frame_id(sg, sg->allows[i].a), b = aa_rx_sg_frame_id(sg, sg->allows[i].b);
        if (a < 0 || b < 0) continue; /* pair names a frame that is not in this graph (e.g. after a sub-scene copy) */
        tmk_bit_put(sg->allowed, sg->wpr, (size_t)a, (size_t)b, sg->allows[i].allowed);
        tmk_bit_put(sg->allowed, sg->wpr, (size_t)b, (size_t)a, sg->allows[i].allowed);
    }
    free(placed);
    sg->serial++;
}

void aa_rx_sg_cl_init(struct aa_rx_sg *sg) {
    if (!sg->inited) tmk_die("aa_rx_sg_cl_init: call aa_rx_sg_init first");
    if (sg->cl_inited) return;
    size_t ng = 0, nm = 0;
    for (size_t i = 0; i < sg->n_def; i++) ng += (size_t)sg->defs[i].n_geom;
    sg->g_frame = (int *)malloc(sizeof(int) * (ng + 1));
    sg->g_shape = (int *)malloc(sizeof(int) * (ng + 1));
    sg->g_mesh = (int *)malloc(sizeof(int) * (ng + 1));
    sg->g_dim = (double *)malloc(sizeof(double) * 3 * (ng + 1));
    sg->meshes = (struct aa_rx_mesh **)malloc(sizeof(void *) * (ng + 1));
    size_t g = 0;
    for (size_t id = 0; id < sg->n_f; id++) {
        tmk_frame_def *d = &sg->defs[sg->order[id]];
        for (int k = 0; k < d->n_geom; k++, g++) {
            sg->g_frame[g] = (int)id;
            sg->g_shape[g] = d->geoms[k]->shape;
            memcpy(sg->g_dim + 3 * g, d->geoms[k]->dim, sizeof(double) * 3);
            sg->g_mesh[g] = -1;
            if (d->geoms[k]->shape == TMK_MESH) {
                size_t m;
                for (m = 0; m < nm; m++)
                    if (sg->meshes[m] == d->geoms[k]->mesh) break;
                if (m == nm) sg->meshes[nm++] = d->geoms[k]->mesh;
                sg->g_mesh[g] = (int)m;
            }
        }
    }
    sg->n_g = ng;
    sg->n_mesh = nm;
    sg->cl_inited = 1;
    sg->serial++;
}

/* ------------------------------------------------------------------ queries */
static void need_init(const struct aa_rx_sg *sg, const char *fn) {
    if (!sg || !sg->inited) tmk_die("%s: scene graph not initialised (aa_rx_sg_init)", fn);
}
size_t aa_rx_sg_config_count(const struct aa_rx_sg *sg) { need_init(sg, "aa_rx_sg_config_count"); return sg->n_q; }
size_t aa_rx_sg_frame_count(const struct aa_rx_sg *sg) { need_init(sg, "aa_rx_sg_frame_count"); return sg->n_f; }

aa_rx_frame_id aa_rx_sg_frame_id(const struct aa_rx_sg *sg, const char *name) {
    need_init(sg, "aa_rx_sg_frame_id");
    if (!name || !*name) return AA_RX_FRAME_ROOT;
    for (size_t id = 0; id < sg->n_f; id++)
        if (0 == strcmp(sg->defs[sg->order[id]].name, name)) return (int)id;
    return AA_RX_FRAME_NONE;
}
const char *aa_rx_sg_frame_name(const struct aa_rx_sg *sg, aa_rx_frame_id id) {
    need_init(sg, "aa_rx_sg_frame_name");
    if (id == AA_RX_FRAME_ROOT) return "";
    if (id < 0 || (size_t)id >= sg->n_f) tmk_die("aa_rx_sg_frame_name: bad id %d", id);
    return sg->defs[sg->order[id]].name;
}
aa_rx_frame_id aa_rx_sg_frame_parent(const struct aa_rx_sg *sg, aa_rx_frame_id id) {
    need_init(sg, "aa_rx_sg_frame_parent");
    if (id < 0 || (size_t)id >= sg->n_f) tmk_die("aa_rx_sg_frame_parent: bad id %d", id);
    return sg->parent[id];
}
aa_rx_config_id aa_rx_sg_config_id(const struct aa_rx_sg *sg, const char *name) {
    need_init(sg, "aa_rx_sg_config_id");
    for (size_t k = 0; k < sg->n_q; k++)
        if (0 == strcmp(sg->config_names[k], name)) return (int)k;
    return AA_RX_CONFIG_NONE;
}
const char *aa_rx_sg_config_name(const struct aa_rx_sg *sg, aa_rx_config_id id) {
    need_init(sg, "aa_rx_sg_config_name");
    if (id < 0 || (size_t)id >= sg->n_q) tmk_die("aa_rx_sg_config_name: bad id %d", id);
    return sg->config_names[id];
}
void aa_rx_sg_config_indices(const struct aa_rx_sg *sg, size_t n, const char **names, aa_rx_config_id *ids) {
    for (size_t i = 0; i < n; i++) ids[i] = aa_rx_sg_config_id(sg, names[i]);
}
void aa_rx_sg_config_names(const struct aa_rx_sg *sg, size_t n, const char **names) {
    need_init(sg, "aa_rx_sg_config_names");
    for (size_t i = 0; i < n && i < sg->n_q; i++) names[i] = sg->config_names[i];
}
void aa_rx_sg_config_set(const struct aa_rx_sg *sg, size_t n_all, size_t n_sub, const aa_rx_config_id *ids,
                         const double *q_sub, double *q_all) {
    (void)sg;
    for (size_t i = 0; i < n_sub; i++) {
        if (ids[i] < 0 || (size_t)ids[i] >= n_all) tmk_die("aa_rx_sg_config_set: bad config id %d", ids[i]);
        q_all[ids[i]] = q_sub[i];
    }
}
int aa_rx_sg_get_limit_pos(const struct aa_rx_sg *sg, aa_rx_config_id id, double *min, double *max) {
    need_init(sg, "aa_rx_sg_get_limit_pos");
    if (id < 0 || (size_t)id >= sg->n_q) tmk_die("aa_rx_sg_get_limit_pos: bad id %d", id);
    if (isnan(sg->lim_min[id])) return -1;
    *min = sg->lim_min[id];
    *max = sg->lim_max[id];
    return 0;
}

/* ------------------------------------------------------------------ frame accessors, chains */
int aa_rx_sg_frame_type(const struct aa_rx_sg *sg, aa_rx_frame_id id) {
    need_init(sg, "aa_rx_sg_frame_type");
    if (id < 0 || (size_t)id >= sg->n_f) tmk_die("aa_rx_sg_frame_type: bad frame id %d", id);
    return sg->type[id];
}
aa_rx_config_id aa_rx_sg_frame_config(const struct aa_rx_sg *sg, aa_rx_frame_id id) {
    need_init(sg, "aa_rx_sg_frame_config");
    if (id < 0 || (size_t)id >= sg->n_f) tmk_die("aa_rx_sg_frame_config: bad frame id %d", id);
    return sg->type[id] == TMK_FIXED ? AA_RX_CONFIG_NONE : sg->qidx[id];
}
void aa_rx_sg_frame_axis(const struct aa_rx_sg *sg, aa_rx_frame_id id, double axis[3]) {
    need_init(sg, "aa_rx_sg_frame_axis");
    if (id < 0 || (size_t)id >= sg->n_f) tmk_die("aa_rx_sg_frame_axis: bad frame id %d", id);
    memcpy(axis, sg->axis + 3 * (size_t)id, sizeof(double) * 3);
}

struct aa_rx_sg_sub {
    const struct aa_rx_sg *sg;
    aa_rx_frame_id root, tip;
    size_t n_sub;
    aa_rx_config_id *ids; /* configuration variables on the chain, root -> tip, each once */
    size_t n_frames;
    aa_rx_frame_id *frames; /* frames on the chain, root -> tip (root itself excluded) */
};

/* reference: tm-blocks.py:202 aa.scene_chain(scene, "", frame) */
struct aa_rx_sg_sub *aa_rx_sg_chain_create(const struct aa_rx_sg *sg, aa_rx_frame_id root, aa_rx_frame_id tip) {
    need_init(sg, "aa_rx_sg_chain_create");
    if (tip < 0 || (size_t)tip >= sg->n_f) tmk_die("aa_rx_sg_chain_create: bad tip frame %d", tip);
    if (root != AA_RX_FRAME_ROOT && (root < 0 || (size_t)root >= sg->n_f)) tmk_die("aa_rx_sg_chain_create: bad root frame %d", root);
    struct aa_rx_sg_sub *s = (struct aa_rx_sg_sub *)calloc(1, sizeof(*s));
    s->sg = sg; s->root = root; s->tip = tip;
    size_t depth = 0;
    int f = tip;
    while (f != root && f >= 0) { depth++; f = sg->parent[f]; }
    if (f != root) tmk_die("aa_rx_sg_chain_create: frame %d is not an ancestor of frame %d", root, tip);
    s->frames = (aa_rx_frame_id *)malloc(sizeof(int) * (depth + 1));
    s->ids = (aa_rx_config_id *)malloc(sizeof(int) * (depth + 1));
    s->n_frames = depth;
    f = tip;
    for (size_t k = depth; k-- > 0;) { s->frames[k] = f; f = sg->parent[f]; }
    for (size_t k = 0; k < depth; k++) {
        int fr = s->frames[k];
        if (sg->type[fr] == TMK_FIXED) continue;
        int seen = 0;
        for (size_t j = 0; j < s->n_sub; j++) seen |= s->ids[j] == sg->qidx[fr];
        if (!seen) s->ids[s->n_sub++] = sg->qidx[fr];
    }
    return s;
}
void aa_rx_sg_sub_destroy(struct aa_rx_sg_sub *s) {
    if (!s) return;
    free(s->frames); free(s->ids); free(s);
}
size_t aa_rx_sg_sub_config_count(const struct aa_rx_sg_sub *s) { return s->n_sub; }
const aa_rx_config_id *aa_rx_sg_sub_configs(const struct aa_rx_sg_sub *s) { return s->ids; }
aa_rx_frame_id aa_rx_sg_sub_tip(const struct aa_rx_sg_sub *s) { return s->tip; }
const struct aa_rx_sg *tmk_sub_sg(const struct aa_rx_sg_sub *s) { return s->sg; }
size_t tmk_sub_frames(const struct aa_rx_sg_sub *s, const aa_rx_frame_id **frames) { *frames = s->frames; return s->n_frames; }

/* ------------------------------------------------------------------ FK entry (device) */
void aa_rx_sg_tf(const struct aa_rx_sg *sg, size_t n_q, const double *q, size_t n_tf, double *TF_rel, size_t ld_rel,
                 double *TF_abs, size_t ld_abs) {
    need_init(sg, "aa_rx_sg_tf");
    if (n_q != sg->n_q) tmk_die("aa_rx_sg_tf: n_q = %zu, scene graph has %zu", n_q, sg->n_q);
    if (n_tf > sg->n_f) tmk_die("aa_rx_sg_tf: n_tf = %zu > frame count %zu", n_tf, sg->n_f);
    if (ld_rel < 7 || ld_abs < 7) tmk_die("aa_rx_sg_tf: leading dimension < 7");
    double *rel = (double *)malloc(sizeof(double) * 7 * (sg->n_f + 1));
    double *ab = (double *)malloc(sizeof(double) * 7 * (sg->n_f + 1));
    tmk_dev_fk64(sg, q, rel, ab);
    for (size_t i = 0; i < n_tf; i++) {
        if (TF_rel) memcpy(TF_rel + i * ld_rel, rel + 7 * i, sizeof(double) * 7);
        if (TF_abs) memcpy(TF_abs + i * ld_abs, ab + 7 * i, sizeof(double) * 7);
    }
    free(rel);
    free(ab);
}

void aa_rx_sg_rel_tf(const struct aa_rx_sg *sg, aa_rx_frame_id parent, aa_rx_frame_id child, const double *TF_abs,
                     size_t ld, double *out) {
    need_init(sg, "aa_rx_sg_rel_tf");
    static const double ident[7] = {0, 0, 0, 1, 0, 0, 0};
    if (child < 0 || (size_t)child >= sg->n_f) tmk_die("aa_rx_sg_rel_tf: bad child id %d", child);
    if (parent != AA_RX_FRAME_ROOT && (parent < 0 || (size_t)parent >= sg->n_f))
        tmk_die("aa_rx_sg_rel_tf: bad parent id %d", parent);
    const double *gp = parent == AA_RX_FRAME_ROOT ? ident : TF_abs + (size_t)parent * ld;
    double inv[7];
    tmk_tfinv(gp, inv);
    tmk_tfmul(inv, TF_abs + (size_t)child * ld, out);
}

/* ------------------------------------------------------------------ mutation */
struct aa_rx_sg *aa_rx_sg_copy(const struct aa_rx_sg *src) {
    struct aa_rx_sg *sg = aa_rx_sg_create();
    sg->n_def = sg->cap_def = src->n_def;
    sg->defs = (tmk_frame_def *)malloc(sizeof(tmk_frame_def) * (src->n_def + 1));
    for (size_t i = 0; i < src->n_def; i++) {
        const tmk_frame_def *s = &src->defs[i];
        tmk_frame_def *d = &sg->defs[i];
        *d = *s;
        d->name = dupstr(s->name);
        d->parent = dupstr(s->parent);
        d->config = s->config ? dupstr(s->config) : NULL;
        d->geoms = (struct aa_rx_geom **)malloc(sizeof(void *) * (size_t)(s->n_geom + 1));
        for (int k = 0; k < s->n_geom; k++) d->geoms[k] = geom_clone(s->geoms[k]);
    }
    for (size_t i = 0; i < src->n_limit; i++) aa_rx_sg_set_limit_pos(sg, src->limits[i].name, src->limits[i].min, src->limits[i].max);
    for (size_t i = 0; i < src->n_allow; i++) allow_name(sg, src->allows[i].a, src->allows[i].b, src->allows[i].allowed);
    sg->inited = 0;
    return sg;
}

void aa_rx_sg_reparent_name(struct aa_rx_sg *sg, const char *new_parent, const char *child, const double *tf7) {
    tmk_frame_def *d = find_def(sg, child);
    if (!d) tmk_die("aa_rx_sg_reparent_name: no frame `%s'", child);
    if (new_parent && *new_parent && !find_def(sg, new_parent))
        tmk_die("aa_rx_sg_reparent_name: no frame `%s'", new_parent);
    if (d->type != TMK_FIXED) tmk_die("aa_rx_sg_reparent_name: `%s' is not a fixed frame", child);
    free(d->parent);
    d->parent = dupstr(new_parent);
    memcpy(d->E, tf7, sizeof(double) * 7);
    double n = sqrt(d->E[0] * d->E[0] + d->E[1] * d->E[1] + d->E[2] * d->E[2] + d->E[3] * d->E[3]);
    for (int i = 0; i < 4; i++) d->E[i] /= n;
    sg->inited = 0;
}

/* allow everything that collides at q: FK + full-report check on the GPU, OR into the scene's set */
void aa_rx_sg_allow_config(struct aa_rx_sg *sg, size_t n_q, const double *q) {
    need_init(sg, "aa_rx_sg_allow_config");
    aa_rx_sg_cl_init(sg);
    size_t n_f = sg->n_f;
    double *ab = (double *)malloc(sizeof(double) * 7 * (n_f + 1));
    aa_rx_sg_tf(sg, n_q, q, n_f, NULL, 7, ab, 7);
    struct aa_rx_cl *cl = aa_rx_cl_create(sg);
    struct aa_rx_cl_set *set = aa_rx_cl_set_create(sg);
    aa_rx_cl_check(cl, n_f, ab, 7, set);
    for (size_t i = 0; i < n_f; i++)
        for (size_t j = 0; j < i; j++)
            if (aa_rx_cl_set_get(set, (int)i, (int)j))
                aa_rx_sg_allow_collision_name(sg, aa_rx_sg_frame_name(sg, (int)i), aa_rx_sg_frame_name(sg, (int)j), 1);
    aa_rx_cl_set_destroy(set);
    aa_rx_cl_destroy(cl);
    free(ab);
}

/* ------------------------------------------------------------------ collision sets */
struct aa_rx_cl_set *aa_rx_cl_set_create(const struct aa_rx_sg *sg) {
    need_init(sg, "aa_rx_cl_set_create");
    struct aa_rx_cl_set *s = (struct aa_rx_cl_set *)calloc(1, sizeof(*s));
    s->n_f = sg->n_f;
    s->wpr = (sg->n_f + 31) / 32;
    s->bits = (uint32_t *)calloc(s->n_f * s->wpr + 1, sizeof(uint32_t));
    return s;
}
void aa_rx_cl_set_destroy(struct aa_rx_cl_set *s) {
    if (!s) return;
    free(s->bits);
    free(s);
}
void aa_rx_cl_set_clear(struct aa_rx_cl_set *s) { memset(s->bits, 0, sizeof(uint32_t) * s->n_f * s->wpr); }
int aa_rx_cl_set_get(const struct aa_rx_cl_set *s, aa_rx_frame_id i, aa_rx_frame_id j) {
    if (i < 0 || j < 0 || (size_t)i >= s->n_f || (size_t)j >= s->n_f) tmk_die("aa_rx_cl_set_get: bad frame id");
    return tmk_bit_get(s->bits, s->wpr, (size_t)i, (size_t)j);
}
void aa_rx_cl_set_set(struct aa_rx_cl_set *s, aa_rx_frame_id i, aa_rx_frame_id j, int v) {
    if (i < 0 || j < 0 || (size_t)i >= s->n_f || (size_t)j >= s->n_f) tmk_die("aa_rx_cl_set_set: bad frame id");
    tmk_bit_put(s->bits, s->wpr, (size_t)i, (size_t)j, v);
    tmk_bit_put(s->bits, s->wpr, (size_t)j, (size_t)i, v);
}
void aa_rx_cl_set_fill(struct aa_rx_cl_set *dst, const struct aa_rx_cl_set *src) {
    if (dst->n_f != src->n_f) tmk_die("aa_rx_cl_set_fill: size mismatch");
    for (size_t k = 0; k < dst->n_f * dst->wpr; k++) dst->bits[k] |= src->bits[k];
}

/* ------------------------------------------------------------------ scene plugins (tmpexec.c:172) */
struct aa_rx_sg *aa_rx_dl_sg(const char *plugin_file, const char *scene_name, struct aa_rx_sg *sg) {
    void *h = dlopen(plugin_file, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "tmkcl: could not open scene plugin `%s': %s\n", plugin_file, dlerror());
        return NULL;
    }
    char sym[512];
    snprintf(sym, sizeof(sym), "aa_rx_dl_sg__%s", scene_name);
    struct aa_rx_sg *(*fn)(struct aa_rx_sg *) = (struct aa_rx_sg * (*)(struct aa_rx_sg *)) dlsym(h, sym);
    if (!fn) {
        fprintf(stderr, "tmkcl: scene plugin `%s' has no scene `%s' (%s)\n", plugin_file, scene_name, sym);
        return NULL;
    }
    if (!sg) sg = aa_rx_sg_create();
    return fn(sg);
}

const char *tmk_version(void) { return "tmkit_b200 0.1 (sm_100a)"; }
