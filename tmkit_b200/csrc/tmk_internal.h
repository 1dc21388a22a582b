/* tmk_internal.h - structures shared by the host C scene graph (sg.c) and the CUDA side (cl.cu).
 * Not installed; the public ABI is include/tmkit_b200/amino_rx.h. */
#ifndef TMK_INTERNAL_H
#define TMK_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include "tmkit_b200/amino_rx.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { TMK_FIXED = 0, TMK_REVOLUTE = 1, TMK_PRISMATIC = 2 };
#define TMK_MP_MAXSUB 24 /* configuration variables of one motion query (= TMK_MAXSUB of the kernels) */
enum { TMK_SPHERE = 0, TMK_BOX = 1, TMK_CYLINDER = 2, TMK_MESH = 3 };

struct aa_rx_mesh {
    int refcount;
    size_t nv;   /* vertices */
    float *verts; /* nv*3 */
    size_t nt;   /* triangles */
    unsigned *idx; /* nt*3 */
    void *compiled; /* BVH built by the first collision context that sees the mesh (owned by cl.cu) */
};
void tmk_mesh_compiled_release(void *compiled);

struct aa_rx_geom {
    int shape;
    double dim[3]; /* sphere: r | box: full extents | cylinder: height, radius */
    struct aa_rx_mesh *mesh;
};

typedef struct {
    char *name;
    char *parent; /* "" = root */
    int type;
    double E[7];
    double axis[3];
    double offset;
    char *config; /* NULL for fixed frames */
    int n_geom;
    struct aa_rx_geom **geoms;
} tmk_frame_def;

typedef struct { char *name; double min, max; } tmk_limit;
typedef struct { char *a, *b; int allowed; } tmk_allow;

struct aa_rx_sg {
    /* definition (mutable until aa_rx_sg_init) */
    size_t n_def, cap_def;
    tmk_frame_def *defs;
    size_t n_limit, cap_limit;
    tmk_limit *limits;
    size_t n_allow, cap_allow;
    tmk_allow *allows;

    /* compiled by aa_rx_sg_init */
    int inited;
    size_t n_f, n_q;
    int *order;      /* sorted id -> def index */
    int *parent;     /* [n_f] */
    int *type;
    double *E;       /* [n_f*7] */
    double *axis;    /* [n_f*3] */
    double *offset;
    int *qidx;       /* [n_f], -1 for fixed */
    char **config_names; /* [n_q] (borrowed from defs) */
    double *lim_min, *lim_max; /* [n_q], NAN when unset */
    uint32_t *allowed; /* n_f x n_f bit matrix, row stride = words_per_row */
    size_t wpr;

    /* compiled by aa_rx_sg_cl_init */
    int cl_inited;
    size_t n_g;
    int *g_frame, *g_shape, *g_mesh;
    double *g_dim;   /* [n_g*3] */
    size_t n_mesh;
    struct aa_rx_mesh **meshes; /* borrowed */

    void *dev; /* device mirror for FK (owned by cl.cu) */
    uint64_t serial; /* bumped by every aa_rx_sg_init / allow change; lets collision contexts notice */
};

struct aa_rx_cl_set {
    size_t n_f, wpr;
    uint32_t *bits;
};

/* bit-matrix helpers */
static inline int tmk_bit_get(const uint32_t *m, size_t wpr, size_t i, size_t j) {
    return (int)((m[i * wpr + (j >> 5)] >> (j & 31)) & 1u);
}
static inline void tmk_bit_put(uint32_t *m, size_t wpr, size_t i, size_t j, int v) {
    if (v) m[i * wpr + (j >> 5)] |= (1u << (j & 31));
    else m[i * wpr + (j >> 5)] &= ~(1u << (j & 31));
}

void tmk_die(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));

/* host-side double transform helpers (scene compilation only; the per-configuration path is on the GPU) */
void tmk_qmul(const double *a, const double *b, double *c);
void tmk_qrot(const double *q, const double *v, double *out);
void tmk_tfmul(const double *a, const double *b, double *c);
void tmk_tfinv(const double *a, double *c);

/* sub-scene-graph internals (sg.c) used by the motion planner (mp.c) */
struct aa_rx_sg_sub;
const struct aa_rx_sg *tmk_sub_sg(const struct aa_rx_sg_sub *s);
size_t tmk_sub_frames(const struct aa_rx_sg_sub *s, const aa_rx_frame_id **frames);

/* implemented in cl.cu */
void tmk_dev_sg_release(void *dev);
void tmk_dev_fk64(const struct aa_rx_sg *sg, const double *q, double *TF_rel, double *TF_abs); /* dense n_f x 7 */
/* cl.cu: all-iterations-on-the-device IK of a parent-linked chain (see ik.cuh) */
int tmk_dev_ik(const struct aa_rx_sg *sg, size_t n_fr, const aa_rx_frame_id *frames, size_t n_sub,
               const aa_rx_config_id *sub_ids, const double *q_all_base, const double *lo, const double *hi,
               const double *goal, size_t n_seed, double *Q, double *err, int max_iter, int enough, int min_iter,
               double tol);

/* cl.cu: the host-buffer calls in two halves (enqueue / wait), for the multi-GPU wrappers in multi.cu */
void tmk_cl_batch_host_begin(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids, size_t n_q,
                             const double *q_base, const void *Q, int layout, size_t ldQ, uint32_t *valid_bits);
void tmk_cl_edges_host_begin(struct aa_rx_cl *cl, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids, size_t n_q,
                             const double *q_base, const double *Q0, const double *Q1, size_t ldQ, double seg_len,
                             uint32_t *valid_bits, int32_t *first_invalid);
size_t tmk_cl_host_end(struct aa_rx_cl *cl, size_t n_units, const uint32_t *valid_bits);
int tmk_cl_device(const struct aa_rx_cl *cl);
void *tmk_cl_stream(const struct aa_rx_cl *cl);
const struct aa_rx_sg *tmk_cl_sg(const struct aa_rx_cl *cl);

#ifdef __cplusplus
}
#endif
#endif
