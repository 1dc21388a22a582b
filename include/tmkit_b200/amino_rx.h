/* amino_rx.h - C ABI of libtmkcl.so, the B200 validity-check library.
 *
 * Drop-in boundary for the one path TMKit drives through Amino: scene-graph forward
 * kinematics (aa_rx_sg_tf) followed by the collision query (aa_rx_cl_check), plus the
 * scene-ingest constructors the reference's generated scene plugins call, plus batched
 * configuration / swept-edge entry points (new).
 *
 * The reference tree holds the CALL SITES of this ABI, not its declarations (those live in
 * Amino's <amino/rx/scenegraph.h>, <amino/rx/scene_collision.h>, <amino/rx/scene_plugin.h>,
 * which are not in /root/reference).  Each entry point cites the reference call site it
 * must satisfy; argument lists are read off those call sites.
 *
 * Conventions (reference: demo/tmpexec.c:319-330,404-409; demo/baxter-sussman/class.robray:61-65):
 *   - scalars are double, ids are int, counts are size_t;
 *   - a transform is 7 doubles: quaternion x,y,z,w then translation x,y,z; arrays of
 *     transforms take an explicit leading dimension (ld >= 7);
 *   - handles are opaque heap objects with create/destroy pairs; numeric buffers are
 *     caller-owned; a scene graph is immutable after aa_rx_sg_init (mutate = copy);
 *   - misuse aborts with a message on stderr (Amino asserts); there is no errno;
 *   - every compute entry point runs on the current CUDA device.  There is NO CPU
 *     fallback: without a usable device the call prints the CUDA error and aborts.
 *   - a handle is used by one thread at a time; distinct handles are independent.  A collision context lives
 *     on the device that was current at aa_rx_cl_create and may have work in flight on ONE stream at a time
 *     (its scratch buffers are per handle): synchronise before switching streams or calling a blocking entry
 *     point after an asynchronous one.
 */
#ifndef TMKIT_B200_AMINO_RX_H
#define TMKIT_B200_AMINO_RX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct aa_rx_sg;
struct aa_rx_cl;
struct aa_rx_cl_set;
struct aa_rx_geom;
struct aa_rx_geom_opt;
struct aa_rx_mesh;

typedef int aa_rx_frame_id;
typedef int aa_rx_config_id;
#define AA_RX_FRAME_ROOT (-1)
#define AA_RX_FRAME_NONE (-2)
#define AA_RX_CONFIG_NONE (-1)

/* ---------------------------------------------------------------------------------------
 * B2: scene ingest - the constructor set the reference's aarxc-generated scene plugins call
 * (Makefile.am:213-231 generates demo/baxter-model.c, demo/sussman-scene.c; tmpexec.c:172 loads them)
 * ------------------------------------------------------------------------------------- */
struct aa_rx_sg *aa_rx_sg_create(void);
void aa_rx_sg_destroy(struct aa_rx_sg *sg); /* demo/tmpexec.c:149 */

/* parent "" (or NULL) = root.  q = quaternion xyzw, v = translation. */
void aa_rx_sg_add_frame_fixed(struct aa_rx_sg *sg, const char *parent, const char *name, const double q[4],
                              const double v[3]);
void aa_rx_sg_add_frame_revolute(struct aa_rx_sg *sg, const char *parent, const char *name, const double q[4],
                                 const double v[3], const char *config_name, const double axis[3], double offset);
void aa_rx_sg_add_frame_prismatic(struct aa_rx_sg *sg, const char *parent, const char *name, const double q[4],
                                  const double v[3], const char *config_name, const double axis[3], double offset);
void aa_rx_sg_set_limit_pos(struct aa_rx_sg *sg, const char *config_name, double min, double max);
/* returns 0 and fills min/max if the configuration variable has a position limit, else -1 */
int aa_rx_sg_get_limit_pos(const struct aa_rx_sg *sg, aa_rx_config_id id, double *min, double *max);

/* Geometry.  opt may be NULL (every geometry attached through this ABI is collision geometry).
 * box: centred, dimension = full extents (tm-blocks.py:213-217).  sphere: centred.
 * cylinder: axis +z, base at the frame origin (Amino convention).
 * mesh: triangle soup, surface only. */
struct aa_rx_geom *aa_rx_geom_box(struct aa_rx_geom_opt *opt, const double dimension[3]);
struct aa_rx_geom *aa_rx_geom_sphere(struct aa_rx_geom_opt *opt, double radius);
struct aa_rx_geom *aa_rx_geom_cylinder(struct aa_rx_geom_opt *opt, double height, double radius);
struct aa_rx_mesh *aa_rx_mesh_create(void);
void aa_rx_mesh_destroy(struct aa_rx_mesh *mesh);
void aa_rx_mesh_set_vertices(struct aa_rx_mesh *mesh, size_t n, const float *vertices, int copy);
void aa_rx_mesh_set_indices(struct aa_rx_mesh *mesh, size_t n, const unsigned *indices, int copy);
struct aa_rx_geom *aa_rx_geom_mesh(struct aa_rx_geom_opt *opt, struct aa_rx_mesh *mesh);
/* attaches (and takes ownership of) geom to the named frame */
void aa_rx_geom_attach(struct aa_rx_sg *sg, const char *frame, struct aa_rx_geom *geom);

/* allow_collision "a" "b";  (demo/baxter-sussman/allowed-collision.robray:1-31) */
void aa_rx_sg_allow_collision_name(struct aa_rx_sg *sg, const char *frame0, const char *frame1, int allowed);
void aa_rx_sg_allow_collision(struct aa_rx_sg *sg, aa_rx_frame_id id0, aa_rx_frame_id id1, int allowed);

/* dlopen `plugin_file`, call its `aa_rx_dl_sg__<scene_name>(sg)` constructor.  sg may be NULL
 * (a new graph is created).  Returns NULL + message on stderr on failure (demo/tmpexec.c:172-177). */
struct aa_rx_sg *aa_rx_dl_sg(const char *plugin_file, const char *scene_name, struct aa_rx_sg *sg);

/* ---------------------------------------------------------------------------------------
 * B1: scene-graph queries and FK
 * ------------------------------------------------------------------------------------- */
void aa_rx_sg_init(struct aa_rx_sg *sg);    /* demo/tmpexec.c:105,336: sort frames, index configs */
void aa_rx_sg_cl_init(struct aa_rx_sg *sg); /* demo/tmpexec.c:106,119,397: prepare collision data */

size_t aa_rx_sg_config_count(const struct aa_rx_sg *sg);                            /* tmpexec.c:114,366 */
size_t aa_rx_sg_frame_count(const struct aa_rx_sg *sg);                             /* tmpexec.c:316,367 */
aa_rx_frame_id aa_rx_sg_frame_id(const struct aa_rx_sg *sg, const char *name);      /* tmpexec.c:328-329 */
const char *aa_rx_sg_frame_name(const struct aa_rx_sg *sg, aa_rx_frame_id id);      /* tmpexec.c:418 */
aa_rx_frame_id aa_rx_sg_frame_parent(const struct aa_rx_sg *sg, aa_rx_frame_id id);
aa_rx_config_id aa_rx_sg_config_id(const struct aa_rx_sg *sg, const char *name);
const char *aa_rx_sg_config_name(const struct aa_rx_sg *sg, aa_rx_config_id id);
void aa_rx_sg_config_indices(const struct aa_rx_sg *sg, size_t n, const char **names, aa_rx_config_id *ids); /* :272 */
void aa_rx_sg_config_names(const struct aa_rx_sg *sg, size_t n, const char **names);                        /* :342 */
/* q_all[ids[i]] = q_sub[i]   (tmpexec.c:289-291, 344) */
void aa_rx_sg_config_set(const struct aa_rx_sg *sg, size_t n_all, size_t n_sub, const aa_rx_config_id *ids,
                         const double *q_sub, double *q_all);

/* FK of one configuration (tmpexec.c:322-325, 407-408).  TF_rel / TF_abs: n_tf transforms with
 * leading dimensions ld_rel / ld_abs.  Runs on the GPU in double precision. */
void aa_rx_sg_tf(const struct aa_rx_sg *sg, size_t n_q, const double *q, size_t n_tf, double *TF_rel, size_t ld_rel,
                 double *TF_abs, size_t ld_abs);
/* pose of child in parent's frame from absolute transforms (tmpexec.c:327-330) */
void aa_rx_sg_rel_tf(const struct aa_rx_sg *sg, aa_rx_frame_id parent, aa_rx_frame_id child, const double *TF_abs,
                     size_t ld_abs, double *tf_rel7);

struct aa_rx_sg *aa_rx_sg_copy(const struct aa_rx_sg *sg); /* tmpexec.c:334 */
/* re-hang `child` under `new_parent` with relative pose tf7; needs aa_rx_sg_init afterwards (tmpexec.c:335-336) */
void aa_rx_sg_reparent_name(struct aa_rx_sg *sg, const char *new_parent, const char *child, const double *tf7);

/* allow every frame pair that collides at q (tmpexec.c:118) */
void aa_rx_sg_allow_config(struct aa_rx_sg *sg, size_t n_q, const double *q);

/* ---------------------------------------------------------------------------------------
 * B1: collision context and collision sets
 * ------------------------------------------------------------------------------------- */
struct aa_rx_cl *aa_rx_cl_create(const struct aa_rx_sg *sg); /* tmpexec.c:398 */
void aa_rx_cl_destroy(struct aa_rx_cl *cl);
void aa_rx_cl_allow(struct aa_rx_cl *cl, aa_rx_frame_id id0, aa_rx_frame_id id1, int allowed);
void aa_rx_cl_allow_set(struct aa_rx_cl *cl, const struct aa_rx_cl_set *set);
void aa_rx_cl_allow_name(struct aa_rx_cl *cl, const char *frame0, const char *frame1, int allowed);

/* Collision query on absolute transforms (tmpexec.c:409).  Returns 0 = free, non-zero = collision.
 * With set != NULL every colliding frame pair is recorded (set is ORed into). */
int aa_rx_cl_check(struct aa_rx_cl *cl, size_t n_tf, const double *TF, size_t ldTF, struct aa_rx_cl_set *set);

struct aa_rx_cl_set *aa_rx_cl_set_create(const struct aa_rx_sg *sg); /* tmpexec.c:399 */
void aa_rx_cl_set_destroy(struct aa_rx_cl_set *set);
void aa_rx_cl_set_clear(struct aa_rx_cl_set *set);                                    /* tmpexec.c:402 */
int aa_rx_cl_set_get(const struct aa_rx_cl_set *set, aa_rx_frame_id i, aa_rx_frame_id j); /* tmpexec.c:415 */
void aa_rx_cl_set_set(struct aa_rx_cl_set *set, aa_rx_frame_id i, aa_rx_frame_id j, int value);
void aa_rx_cl_set_fill(struct aa_rx_cl_set *dst, const struct aa_rx_cl_set *src);     /* dst |= src */

/* ---------------------------------------------------------------------------------------
 * Batched entry points (new; same conventions).  A batch varies the n_sub configuration
 * variables sub_ids[]; all others stay at q_all_base (what Amino's state validity checker
 * does per state: scatter q_sub into q_all, FK, check - SURVEY.md 3.1).
 * valid_bits: bit (c & 31) of word c >> 5 is 1 iff configuration / edge c is collision-free.
 * ------------------------------------------------------------------------------------- */

/* layouts of a batch of sub-configurations */
#define TMK_Q_ROWS_F64 0 /* double Q[n][ld]   (row per configuration; the Amino / OMPL-facing layout) */
#define TMK_Q_ROWS_F32 1 /* float  Q[n][ld]   */
#define TMK_Q_SOA_F32 2  /* float  Q[n_sub][ld] (ld >= n; joint-major, coalesced) */

/* FK for a batch: TF_abs_out[c][k] (7 doubles each, dense) = absolute transform of frames[k]
 * (frames == NULL: all frames).  Q: TMK_Q_ROWS_F64 host buffer. */
void aa_rx_sg_tf_batch(const struct aa_rx_sg *sg, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                       size_t n_q, const double *q_all_base, const double *Q_sub, size_t ldQ, size_t n_frames,
                       const aa_rx_frame_id *frames, double *TF_abs_out);

/* The same on device buffers, asynchronous on `stream` (NULL = the default stream).  dQ: the sub-configurations in one of
 * the TMK_Q_* layouts; dTF: TMK_TF_ROWS_F64 = double TF[c][k][7] (the host layout above; ld_tf unused) or
 * TMK_TF_SOA_F32 = float component planes TF[(k * 7 + i) * ld_tf + c], ld_tf >= n_cfg (coalesced: the layout to keep
 * transforms in on the device). */
#define TMK_TF_ROWS_F64 0
#define TMK_TF_SOA_F32 1
void aa_rx_sg_tf_batch_dev(const struct aa_rx_sg *sg, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                           size_t n_q, const double *q_all_base, const void *dQ, int layout, size_t ld, size_t n_frames,
                           const aa_rx_frame_id *frames, void *dTF, int tf_layout, size_t ld_tf, void *stream);

/* Validity of n_cfg configurations; host buffers; blocking.  Returns the number of colliding
 * configurations. */
size_t aa_rx_cl_check_batch(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                            size_t n_q, const double *q_all_base, const double *Q_sub, size_t ldQ,
                            uint32_t *valid_bits);
size_t aa_rx_cl_check_batch_f32(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                size_t n_q, const double *q_all_base, const float *Q_sub, size_t ldQ,
                                uint32_t *valid_bits);

/* Swept edges, OMPL DiscreteMotionValidator semantics: edge e is valid iff Q1[e]
 * and the interior states k/nd (k = 1..nd-1, nd = ceil(|Q1-Q0|_2 / seg_len)) are all valid; Q0[e]
 * is assumed valid.  first_invalid (optional): position in OMPL's test order (0 = Q1, then the
 * FIFO bisection of 1..nd-1) of the first invalid state, -1 for a valid edge.  Returns the
 * number of invalid edges. */
size_t aa_rx_cl_check_edges(struct aa_rx_cl *cl, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                            size_t n_q, const double *q_all_base, const double *Q0, const double *Q1, size_t ldQ,
                            double seg_len, uint32_t *valid_bits, int32_t *first_invalid);

/* Device-resident variants: dQ* and d_valid_bits are device pointers on the current device;
 * work is enqueued on `stream` (a cudaStream_t, NULL = the handle's own stream) and NOT
 * synchronised.  layout is one of TMK_Q_*.  q_all_base is a host buffer (read at call time). */
void aa_rx_cl_check_batch_dev(struct aa_rx_cl *cl, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                              size_t n_q, const double *q_all_base, const void *dQ, int layout, size_t ld,
                              uint32_t *d_valid_bits, void *stream);
void aa_rx_cl_check_edges_dev(struct aa_rx_cl *cl, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                              size_t n_q, const double *q_all_base, const void *dQ0, const void *dQ1, int layout,
                              size_t ld, double seg_len, uint32_t *d_valid_bits, int32_t *d_first_invalid,
                              void *stream);
void aa_rx_cl_sync(struct aa_rx_cl *cl, void *stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU (new; SURVEY.md 8e, BASELINE.json configs[4]).  The path shards: every GPU holds a replica of
 * the scene image and checks one contiguous slice of the batch; slices are multiples of TMK_SHARD_ALIGN
 * units, so every GPU owns whole words of the validity bitmap.  Only the bitmap is exchanged.
 * The reference's caller is one single-threaded planner with one handle (lisp/python/tmsmtpy.lisp:55-60,
 * demo/tmpexec.c:397-399): struct aa_rx_cl_multi is that one handle over all GPUs of the box.
 * ------------------------------------------------------------------------------------- */
#define TMK_SHARD_ALIGN 1024
#define TMK_SHARD_ID_BYTES 128
/* slice of rank `rank` among `world`: units [lo, hi) of [0, n); tmk_shard_words = words per rank in the gathered
 * bitmap (the gathered bitmap has world * tmk_shard_words(n, world) words; the first ceil(n/32) are the result) */
size_t tmk_shard_per_rank(size_t n, int world);
void tmk_shard_bounds(size_t n, int world, int rank, size_t *lo, size_t *hi);
size_t tmk_shard_words(size_t n, int world);

/* (1) one process, several devices.  devices == NULL: 0 .. n_dev-1; n_dev <= 0: all visible devices.
 * The per-device contexts are ordinary aa_rx_cl handles (aa_rx_cl_multi_context) living on their device. */
struct aa_rx_cl_multi;
struct aa_rx_cl_multi *aa_rx_cl_multi_create(const struct aa_rx_sg *sg, int n_dev, const int *devices);
void aa_rx_cl_multi_destroy(struct aa_rx_cl_multi *m);
int aa_rx_cl_multi_device_count(const struct aa_rx_cl_multi *m);
int aa_rx_cl_multi_device(const struct aa_rx_cl_multi *m, int k);
struct aa_rx_cl *aa_rx_cl_multi_context(struct aa_rx_cl_multi *m, int k);
void aa_rx_cl_multi_allow(struct aa_rx_cl_multi *m, aa_rx_frame_id id0, aa_rx_frame_id id1, int allowed);
void aa_rx_cl_multi_allow_set(struct aa_rx_cl_multi *m, const struct aa_rx_cl_set *set);
void aa_rx_cl_multi_allow_name(struct aa_rx_cl_multi *m, const char *frame0, const char *frame1, int allowed);
void aa_rx_cl_multi_track_collisions(struct aa_rx_cl_multi *m, int enable);
void aa_rx_cl_multi_batch_collisions(struct aa_rx_cl_multi *m, struct aa_rx_cl_set *set);
/* host buffers, blocking; same arguments and results as the single-device calls.  Every device copies its
 * slice of the rows in and its words straight into valid_bits (no collective). */
size_t aa_rx_cl_check_batch_multi(struct aa_rx_cl_multi *m, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                  size_t n_q, const double *q_all_base, const double *Q_sub, size_t ldQ,
                                  uint32_t *valid_bits);
size_t aa_rx_cl_check_batch_multi_f32(struct aa_rx_cl_multi *m, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                      size_t n_q, const double *q_all_base, const float *Q_sub, size_t ldQ,
                                      uint32_t *valid_bits);
size_t aa_rx_cl_check_edges_multi(struct aa_rx_cl_multi *m, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                                  size_t n_q, const double *q_all_base, const double *Q0, const double *Q1, size_t ldQ,
                                  double seg_len, uint32_t *valid_bits, int32_t *first_invalid);
/* device-resident, asynchronous: dQ[k] = device k's slice (rows lo_k .. hi_k-1, `layout` / `ld` as in the
 * single-device call), d_bits_all[k] = device k's copy of the gathered bitmap.  Kernels run on every context's
 * own stream, then ONE in-place ncclAllGather (communicators from ncclCommInitAll) completes every copy. */
void aa_rx_cl_check_batch_multi_dev(struct aa_rx_cl_multi *m, size_t n_cfg, size_t n_sub, const aa_rx_config_id *sub_ids,
                                    size_t n_q, const double *q_all_base, const void *const *dQ, int layout, size_t ld,
                                    uint32_t *const *d_bits_all);
void aa_rx_cl_check_edges_multi_dev(struct aa_rx_cl_multi *m, size_t n_edge, size_t n_sub, const aa_rx_config_id *sub_ids,
                                    size_t n_q, const double *q_all_base, const void *const *dQ0, const void *const *dQ1,
                                    int layout, size_t ld, double seg_len, uint32_t *const *d_bits_all,
                                    int32_t *const *d_first_slices);
void aa_rx_cl_multi_sync(struct aa_rx_cl_multi *m);

/* (2) one process per GPU (torchrun / MPI style launches): rank 0 makes an id, the caller distributes its
 * TMK_SHARD_ID_BYTES bytes, every rank builds the communicator on its current device.  The call checks this
 * rank's slice (dQ*_slice) into slot `rank` of d_bits_all and all-gathers in place, on `stream` (default) or,
 * with tmk_shard_comm_overlap(c, 1), on the communicator's own high-priority stream so that the gather of one
 * batch runs beside the kernels of the next (tmk_shard_comm_join makes a stream wait for what is in flight). */
struct tmk_shard_comm;
void tmk_shard_unique_id(void *id_bytes);
struct tmk_shard_comm *tmk_shard_comm_create(int rank, int world, const void *id_bytes);
void tmk_shard_comm_destroy(struct tmk_shard_comm *c);
int tmk_shard_comm_rank(const struct tmk_shard_comm *c);
int tmk_shard_comm_world(const struct tmk_shard_comm *c);
void tmk_shard_comm_overlap(struct tmk_shard_comm *c, int enable);
void tmk_shard_comm_join(struct tmk_shard_comm *c, void *stream);
void aa_rx_cl_check_batch_shard_dev(struct aa_rx_cl *cl, struct tmk_shard_comm *c, size_t n_total, size_t n_sub,
                                    const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                    const void *dQ_slice, int layout, size_t ld, uint32_t *d_bits_all, void *stream);
void aa_rx_cl_check_edges_shard_dev(struct aa_rx_cl *cl, struct tmk_shard_comm *c, size_t n_total, size_t n_sub,
                                    const aa_rx_config_id *sub_ids, size_t n_q, const double *q_all_base,
                                    const void *dQ0_slice, const void *dQ1_slice, int layout, size_t ld, double seg_len,
                                    uint32_t *d_bits_all, int32_t *d_first_slice, void *stream);

/* Pinned host memory for the callers of the host-buffer entry points: pageable rows are staged by the driver
 * at a fraction of the PCIe rate.  aa_rx_host_register returns 0 on success, -1 otherwise. */
void *aa_rx_host_alloc(size_t bytes);
void aa_rx_host_free(void *p);
int aa_rx_host_register(void *p, size_t bytes);
void aa_rx_host_unregister(void *p);

/* OR of the colliding frame pairs over the last batch (the :track-collisions feed-back,
 * lisp/python/tmsmtpy.lisp:59; consumed by lisp/itmp-rec.lisp:230-247).  Enabled per handle. */
void aa_rx_cl_track_collisions(struct aa_rx_cl *cl, int enable);
void aa_rx_cl_batch_collisions(struct aa_rx_cl *cl, struct aa_rx_cl_set *set);

/* Counters of the last batch (for the roofline's algorithmic-flop figure). */
struct tmk_batch_stats {
    uint64_t n_units;          /* configurations or edges */
    uint64_t n_states;         /* configurations evaluated (edges: states actually tested) */
    uint64_t n_joints;         /* device FK joints per state */
    uint64_t n_moving_geoms;   /* moving collision geometries per state */
    uint64_t n_pairs;          /* candidate pairs per state (after allowed mask, static pairs hoisted) */
    uint64_t n_narrow[8];      /* narrowphase calls by class: 0 sphere-x, 1 box-box, 2 gjk, 3 mesh leaf */
    uint64_t n_uncertain;      /* states re-evaluated in double precision */
    uint64_t n_invalid;        /* colliding units */
    uint64_t n_launches;       /* kernels launched */
    double gpu_ms;             /* device time of the last batch when stats timing is on, else 0 */
};
void aa_rx_cl_batch_stats(struct aa_rx_cl *cl, struct tmk_batch_stats *out);
void aa_rx_cl_stats_enable(struct aa_rx_cl *cl, int enable); /* instrumented (slower) runs */

/* ---------------------------------------------------------------------------------------
 * "Next" rows: the caller of the path.  Host C on top of the batched entry points
 * above; the shapes follow what robray:motion-plan drives in Amino (lisp/python/tmsmtpy.lisp:55-60:
 * sub-scene-graph, start configuration, workspace goal, :simplify t, :track-collisions t, :timeout).
 * ------------------------------------------------------------------------------------- */
int aa_rx_sg_frame_type(const struct aa_rx_sg *sg, aa_rx_frame_id id);   /* 0 fixed, 1 revolute, 2 prismatic */
aa_rx_config_id aa_rx_sg_frame_config(const struct aa_rx_sg *sg, aa_rx_frame_id id); /* AA_RX_CONFIG_NONE if fixed */
void aa_rx_sg_frame_axis(const struct aa_rx_sg *sg, aa_rx_frame_id id, double axis[3]);

/* chain root -> tip (demo/domains/blocksworld/tm-blocks.py:202: aa.scene_chain(scene, "", frame));
 * root may be AA_RX_FRAME_ROOT.  The sub-scene-graph lists the configuration variables on the chain. */
struct aa_rx_sg_sub;
struct aa_rx_sg_sub *aa_rx_sg_chain_create(const struct aa_rx_sg *sg, aa_rx_frame_id root, aa_rx_frame_id tip);
void aa_rx_sg_sub_destroy(struct aa_rx_sg_sub *ssg);
size_t aa_rx_sg_sub_config_count(const struct aa_rx_sg_sub *ssg);
const aa_rx_config_id *aa_rx_sg_sub_configs(const struct aa_rx_sg_sub *ssg);
aa_rx_frame_id aa_rx_sg_sub_tip(const struct aa_rx_sg_sub *ssg);

/* One motion query: RRT-Connect over the sub-scene-graph's joints (bounds = joint limits, range =
 * 20 % and edge resolution = 1 % of the state-space extent: the OMPL defaults) with
 * every extension checked through aa_rx_cl_check_edges, many extensions per call. */
struct aa_rx_mp;
struct aa_rx_mp *aa_rx_mp_create(const struct aa_rx_sg_sub *ssg);
void aa_rx_mp_destroy(struct aa_rx_mp *mp);
/* start configuration (all n_all variables); collisions present at the start are allowed, as Amino does */
void aa_rx_mp_set_start(struct aa_rx_mp *mp, size_t n_all, const double *q_all);
/* workspace goal for the tip frame (7 doubles): inverse kinematics from several seeds, solutions that
 * violate limits or collide are dropped.  Returns the number of joint-space goals found (0 = failure). */
int aa_rx_mp_set_wsgoal(struct aa_rx_mp *mp, const double *E7);
/* joint-space goal (n_sub values); returns 1 if it is valid */
int aa_rx_mp_set_goal(struct aa_rx_mp *mp, size_t n_sub, const double *q_sub);
void aa_rx_mp_set_simplify(struct aa_rx_mp *mp, int simplify);
void aa_rx_mp_set_track_collisions(struct aa_rx_mp *mp, int track);
void aa_rx_mp_set_seed(struct aa_rx_mp *mp, uint64_t seed);
/* Plan.  On success returns 0 and a malloc'ed path of *n_path rows x n_all doubles (caller frees);
 * on failure (timeout in seconds reached, no goal) returns -1 (-> planning-failure, lisp/tm-plan.lisp:3-7). */
int aa_rx_mp_plan(struct aa_rx_mp *mp, double timeout, size_t *n_path, double **path_all);
/* frame pairs that blocked extensions of the last query (the :collision constraint feed-back) */
void aa_rx_mp_collisions(struct aa_rx_mp *mp, struct aa_rx_cl_set *set);
struct tmk_mp_stats {
    uint64_t n_iterations, n_edge_calls, n_edges_checked, n_nodes, n_ik_iterations, n_goals;
    double ik_s, rrt_s, simplify_s;
};
void aa_rx_mp_stats(const struct aa_rx_mp *mp, struct tmk_mp_stats *out);
/* measurement aid: record the end points (n x n_sub each) of every edge the query checks, so that a
 * CPU baseline can be timed on exactly the same work */
void aa_rx_mp_edge_log(struct aa_rx_mp *mp, int enable);
size_t aa_rx_mp_edge_log_get(const struct aa_rx_mp *mp, const double **q0, const double **q1);

/* Plan files (doc/md/planfile.md:13-76; grammar src/planfile.l:77-192): line-based text,
 * "a ACTION args..", "r frame new_parent", "m config names..", "p values..", "# comment". */
struct tmk_plan;
struct tmk_plan *tmk_plan_create(void);
void tmk_plan_destroy(struct tmk_plan *p);
void tmk_plan_add_action(struct tmk_plan *p, const char *text);
void tmk_plan_add_reparent(struct tmk_plan *p, const char *frame, const char *new_parent);
void tmk_plan_add_motion(struct tmk_plan *p, size_t n_names, const char *const *names, size_t n_points, const double *points,
                         size_t ld);
/* 0 on success; -1 + message on stderr otherwise (src/tmplan.c:301-305 convention) */
int tmk_plan_write(const struct tmk_plan *p, const char *path);
struct tmk_plan *tmk_plan_parse(const char *path);
size_t tmk_plan_op_count(const struct tmk_plan *p);
/* kind: 'a', 'r' or 'm' */
int tmk_plan_op_kind(const struct tmk_plan *p, size_t i);
/* 'a': the action text | 'r': k = 0 frame, 1 new parent | 'm': the k-th configuration name */
const char *tmk_plan_op_text(const struct tmk_plan *p, size_t i, size_t k);
size_t tmk_plan_motion_names(const struct tmk_plan *p, size_t i);
size_t tmk_plan_motion_points(const struct tmk_plan *p, size_t i);
const double *tmk_plan_motion_data(const struct tmk_plan *p, size_t i); /* points x names, dense */

/* Device-time instrumentation for bench.py's roofline: while enabled, CUDA events are recorded on
 * the launching stream around the main validity kernel and around the double-precision re-check
 * kernel of every batch.  aa_rx_cl_kernel_time synchronises those events, returns the number of
 * timed batches, writes the summed device times (ms) and clears the record. */
void aa_rx_cl_kernel_timing(struct aa_rx_cl *cl, int enable);
size_t aa_rx_cl_kernel_time(struct aa_rx_cl *cl, double *ms_main, double *ms_recheck);

/* measured FP32 (FFMA) throughput of the current device in TFLOP/s - the FP32 roofline's denominator */
double tmk_fp32_peak_tflops(void);

const char *tmk_version(void);

#ifdef __cplusplus
}
#endif
#endif
