/* ref_check_state.c - TEST INFRASTRUCTURE, built only where a real Amino installation exists (make -C oracle ref).
 *
 * The validity check of the reference's plan executor, /root/reference/demo/tmpexec.c:362-428 (check_state:
 * aa_rx_sg_tf followed by aa_rx_cl_check), restated as a loop over a file of configurations and linked against the
 * REAL libamino / libamino-collision (FCL underneath).  Its verdicts are what oracle/tmk_oracle.c restates; when this
 * program can be built, oracle/ref_compare.py diffs the two and the parity of the repository becomes pinned.  Amino is
 * not in /root/reference and is not installed in the build image, so by default nothing here is compiled and
 * bench.py reports `amino_fcl: absent`.  Only Amino's public API is used; no reference source is copied.
 *
 *   check_state_ref <scene plugin .so> <scene name> <configs.bin> <verdicts.bin>
 *     configs.bin : uint64 n, uint64 n_q, then n x n_q doubles (scene-graph configuration order)
 *     verdicts.bin: n bytes (1 = collision) ; stdout: seconds per check
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <amino.h>
#include <amino/rx/rxtype.h>
#include <amino/rx/scenegraph.h>
#include <amino/rx/scene_collision.h>
#include <amino/rx/scene_plugin.h>

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s plugin.so scene configs.bin verdicts.bin\n", argv[0]); return 2; }
    struct aa_rx_sg *sg = aa_rx_dl_sg(argv[1], argv[2], NULL);       /* tmpexec.c:172 */
    if (!sg) return 2;
    aa_rx_sg_init(sg);                                               /* tmpexec.c:105-106 */
    aa_rx_sg_cl_init(sg);
    FILE *in = fopen(argv[3], "rb");
    uint64_t n = 0, n_q = 0;
    if (!in || fread(&n, 8, 1, in) != 1 || fread(&n_q, 8, 1, in) != 1) return 2;
    if (n_q != aa_rx_sg_config_count(sg)) { fprintf(stderr, "n_q mismatch\n"); return 2; }
    double *Q = (double *)malloc(sizeof(double) * n * n_q);
    if (fread(Q, sizeof(double), n * n_q, in) != n * n_q) return 2;
    fclose(in);
    aa_rx_sg_allow_config(sg, n_q, Q);                               /* tmpexec.c:118: row 0 is the start configuration */
    aa_rx_sg_cl_init(sg);
    struct aa_rx_cl *cl = aa_rx_cl_create(sg);                       /* tmpexec.c:398 */
    size_t n_f = aa_rx_sg_frame_count(sg);
    double *TF_rel = (double *)malloc(sizeof(double) * 7 * n_f), *TF_abs = (double *)malloc(sizeof(double) * 7 * n_f);
    uint8_t *out = (uint8_t *)malloc(n);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint64_t c = 0; c < n; c++) {
        aa_rx_sg_tf(sg, n_q, Q + c * n_q, n_f, TF_rel, 7, TF_abs, 7); /* tmpexec.c:407-408 */
        out[c] = aa_rx_cl_check(cl, n_f, TF_abs, 7, NULL) ? 1 : 0;    /* tmpexec.c:409, early-out form */
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    FILE *o = fopen(argv[4], "wb");
    fwrite(out, 1, n, o);
    fclose(o);
    printf("%.9g\n", ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / (double)(n ? n : 1));
    return 0;
}
