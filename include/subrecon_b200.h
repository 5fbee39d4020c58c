/*
 * subrecon_b200.h — C ABI of the B200-native SubRecon hot path (libsubrecon_b200.so).
 *
 * The reference (chrismonit/SubRecon) has no FFI: its seam is the Java type Callable<SiteResult>
 * — one `JointBranchReconstruction` task per alignment column submitted to a fixed thread pool and
 * joined in column order (/root/reference/src/subrecon/SubRecon.java:109-142). This library
 * replaces exactly that loop: the Java host flattens what it already holds after `init()`
 * (SubRecon.java:157-268) into plain arrays, makes ONE batched call, and builds its `SiteResult`s
 * (recon/SiteResult.java:91-93) from the returned arrays. INTEGRATION.md shows the JNI stub.
 *
 * Conventions: every function returns 0 on success or a negative sr_status; the message is
 * available from sr_last_error(). Nothing throws or aborts across the ABI. There is NO CPU
 * fallback: without a CUDA device sr_create fails with SR_ERR_CUDA. A context is bound to one
 * GPU and is not thread-safe (one context per caller thread; one per GPU for multi-GPU sharding).
 * sr_set_* copy their arguments; the caller keeps ownership of every buffer it passes.
 */
#ifndef SUBRECON_B200_H
#define SUBRECON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR_N_STATES 20          /* amino acids, PAL order A R N D C Q E G H I L K M F P S T W Y V */
#define SR_STATE_MISSING 255    /* any state code outside 0..19 is missing data
                                   (recon/JointBranchReconstruction.java:199-205) */

typedef enum {
    SR_OK = 0,
    SR_ERR_ARG = -1,            /* bad argument (the reference: ParameterException, SubRecon.java:200-220) */
    SR_ERR_CUDA = -2,           /* CUDA runtime error / no device */
    SR_ERR_OOM = -3,            /* host or device allocation failed */
    SR_ERR_STATE = -4,          /* call order: model / rates / tree / alignment not set yet */
    SR_ERR_UNSUPPORTED = -5
} sr_status;

typedef struct sr_ctx sr_ctx;

/* One compacted column: what SiteResult's constructor scan extracts from the 400 probabilities
 * (recon/SiteResult.java:69-83) plus the print decision of SubRecon.java:138. */
#define SR_COMPACT_MAX 8
typedef struct {
    double  lnl;                    /* marginal site log-likelihood (SiteResult.getMarginalLnL)        */
    double  max_ii_prob;            /* largest diagonal entry   (SiteResult.getMaxIIProb)             */
    double  max_prob;               /* largest entry            (SiteResult.getMaxProb)               */
    int32_t n_kept;                 /* number of entries >= threshold (may exceed SR_COMPACT_MAX)      */
    int32_t interesting;            /* maxII <= 1-threshold && maxProb >= threshold (SubRecon.java:138) */
    int16_t pair[SR_COMPACT_MAX];   /* a*20+b of the first SR_COMPACT_MAX kept entries, canonical order */
    double  prob[SR_COMPACT_MAX];
} sr_compact;

/* Per-kernel device timings accumulated since the last sr_profile_reset (CUDA events on the
 * launching stream). ms_* are sums over launches. */
typedef struct {
    double  ms_pbuild;  int64_t n_pbuild;       /* K1 + K1b: matrix and clade-table builders (launches of both) */
    double  ms_prune;   int64_t n_prune;        /* K2: post-order pruning (dominant)  */
    double  ms_root;    int64_t n_root;         /* K3: root-branch joint / marginal   */
    int64_t columns_pruned;                     /* columns processed by K2 launches   */
    int64_t columns_in, columns_unique;         /* "compress_patterns": columns seen / distinct patterns computed */
} sr_profile;

/* Replaces: `Executors.newFixedThreadPool(nThreads)` (SubRecon.java:109) — the worker pool becomes a GPU. */
int sr_create(sr_ctx **out, int device);
void sr_destroy(sr_ctx *ctx);
const char *sr_last_error(const sr_ctx *ctx);   /* ctx may be NULL: error of the failed sr_create */

/* Replaces the model argument of the task constructor (recon/JointBranchReconstruction.java:49-56):
 * the eigensystem PAL holds in MatrixExponential.{Eval,Evec,Ievc} after any setDistance call
 * (row-major, Evec[i][k], Ievc[k][j]) and SubRecon's `pi` (SubRecon.java:222). */
int sr_set_model(sr_ctx *ctx, const double eval[20], const double evec[400], const double ievc[400],
                 const double pi[20]);

/* Replaces `RateDistribution.getRate(i)/getNumberOfRates()` (JointBranchReconstruction.java:86-87).
 * Class probabilities are 1/K as the reference assumes (it divides by K, :71,:138). */
int sr_set_rates(sr_ctx *ctx, int K, const double *rates);

/* Replaces the `Tree` argument: CSR children in Newick order, branch length to the parent per node,
 * alignment row per tip (-1 for internal nodes). The root must have exactly two children
 * (SubRecon.java:212-213, JointBranchReconstruction.java:59-60); polytomies are allowed below it. */
int sr_set_tree(sr_ctx *ctx, int32_t n_nodes, const int32_t *child_off, const int32_t *child_idx,
                const double *blen, const int32_t *leaf_row, int32_t root);

/* Replaces `alignment.getStateBySequenceName(name, site)` (molevo/AdvancedAlignmentAminoAcid.java:34-40):
 * states[taxon*pitch + site], 0..19 or SR_STATE_MISSING, uploaded once and kept in HBM.
 * With option "input_chars" = 1 the bytes are the alignment's residue LETTERS instead (what PAL's
 * `SimpleAlignment.getData(seq, site)` returns, Latin-1) and the device translates them exactly as
 * `AminoAcids::getStateImpl` does (either case, PAL order; gaps, '?', '*', B J O U X Z and any other byte are
 * missing data) — the host then needs no per-cell encode loop at all. Applies to sr_run_streamed as well. */
int sr_set_alignment(sr_ctx *ctx, int32_t n_taxa, int64_t n_sites, const uint8_t *states, int64_t pitch);

/* Replaces the submit/join loop (SubRecon.java:112-142) for columns [site0, site0+n) of the resident
 * alignment. lnl[n]; joint[n*400] (a*20+b, a = state at root child 0, b = at root child 1) or NULL;
 * compact[n] or NULL (filled with `threshold`). Host buffers; pinned ones (sr_host_alloc) are DMA'd directly. */
int sr_run(sr_ctx *ctx, int64_t site0, int64_t n, double *lnl, double *joint, sr_compact *compact,
           double threshold);

/* Same computation with DEVICE output pointers, enqueued on `stream` (a cudaStream_t, may be NULL = the
 * context's own stream) and not synchronised: for callers that keep results in HBM. */
int sr_run_device(sr_ctx *ctx, int64_t site0, int64_t n, double *d_lnl, double *d_joint, sr_compact *d_compact,
                  double threshold, void *stream);

/* End-to-end variant: the alignment stays in HOST memory (states[taxon*pitch + site]); column chunks are
 * uploaded, computed and downloaded in a double-buffered pipeline. Nothing needs to be resident. */
int sr_run_streamed(sr_ctx *ctx, int32_t n_taxa, int64_t n_sites, const uint8_t *states, int64_t pitch,
                    int64_t site0, int64_t n, double *lnl, double *joint, sr_compact *compact, double threshold);

/* Pinned host memory for the buffers above (JNI: wrap in a direct ByteBuffer). */
void *sr_host_alloc(size_t bytes);
void sr_host_free(void *p);

/* Tuning / introspection. Options: "tile_m" (MMA column tiles per warp: 2 or 4), "warps" (consumer warps per CTA: 12 =
 * default, on re-allocated registers; 8, 11, 15 = tuning variants), "realloc" (0/1: producer warp on re-allocated
 * registers (setmaxnreg), 12 x tile_m 2 or 8 x tile_m 4 consumer warps), "holes" (pauses inside a warp's MMA burst: 0, 1, 2
 * or the digits of the k-steps to pause before), "rotate" (0/1 burst hand-off between the warps of an SM sub-partition),
 * "compress_patterns" (0/1: compute each distinct column pattern of a batch once and copy the result to its
 * duplicates; off by default), "chunk_cols" / "stream_chunk_cols" (columns per launch / per pipeline stage),
 * "stream_taper" (0/1: short pipeline chunks at both ends of a streamed run, on by default), "input_chars" (0/1:
 * alignment bytes are residue letters, see sr_set_alignment), "cherry_tables" (0/1, default 1: tabulate every cherry's
 * contribution per state pair once per rate class and gather it instead of contracting it per column) with
 * "cherry_table_mb" (memory cap, default 4096), "cherry_pairs" (0/1, default 1: a step may gather two such clades) and
 * "cherry_pitchforks" (0/1, default 1: the same one level up, a cherry plus a tip, 9261 state triples), "profile"
 * (0/1: record per-kernel events), "rescale_every" (edges between power-of-two rescalings, default 8). */
int sr_set_option(sr_ctx *ctx, const char *name, int64_t value);
int sr_get_profile(sr_ctx *ctx, sr_profile *out);     /* synchronises the context's streams */
int sr_profile_reset(sr_ctx *ctx);
/* Schedule facts for the current tree: number of pruning steps, stack depth needed, and the
 * algorithmic flops per column per rate class (820*E_int + 20*E_leaf + 1200). */
int sr_schedule_info(sr_ctx *ctx, int64_t *n_steps, int32_t *stack_depth, double *flops_per_column_rate);
/* Host-only (no GPU needed): the fused pruning schedule sr_set_tree would build. steps[12*i..] = {word, alignment row of
 * gather 0, 1, 2; second row of the step's first and second clade gather, their third rows; first table row of each, 0, 0};
 * word bits: 0-1 MV (1 = A=P·A,
 * 2 = A=pop∘P·A), 2-3 number of tip gathers before the MV, 4 first gather starts a new partial, 5 one tip gather after
 * the MV, 6 rescale, 7 push, 8/9 store as root child A/B, 10-11 and 12-13 positions (1-based, among the step's gathers) of
 * up to two collapsed-clade gathers, 14 / 15 the first / second of them is a pitchfork (three tips).
 * node_slot[v] = position of node v's parent edge in the matrix stream (per step: MV first, then tips), -1 if it has
 * none, -2-c if v is the top of collapsed clade c (a cherry: two tip children; or a pitchfork: a cherry and a tip — with
 * `cherry_tables` its contribution to its parent is tabulated per state combination and gathered like a tip's).
 * This replaces the recursion order of downTreeMarginal (JointBranchReconstruction.java:191-247). */
int sr_schedule_dump(int32_t n_nodes, const int32_t *child_off, const int32_t *child_idx, const double *blen,
                     const int32_t *leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables, int64_t cap_steps,
                     int32_t *steps, int32_t *node_slot, int64_t *n_steps, int32_t *stack_depth);
/* Live FP64 roofline probe: runs a register-resident DMMA (mma.sync m8n8k4 f64) loop for ~`ms` and
 * returns TFLOP/s — the denominator bench.py reports the pruning kernel against. */
int sr_measure_fp64_peak(sr_ctx *ctx, double ms, double *tflops);
/* Copies the device-built transition matrix of (node, rate class) to host, row-major P[i][j]
 * (parity hook for `model.getTransitionProbabilities`, JointBranchReconstruction.java:217);
 * node == root returns the merged root-branch matrix (:99). */
int sr_debug_pmatrix(sr_ctx *ctx, int32_t node, int32_t rate_class, double out[400]);
/* Copies the device-built table of a collapsed clade (node = its top node) to host, 20 doubles per row; *n_rows = 441
 * for a cherry (tips a, b): out[(s_a*21 + s_b)*20 + i] = Σ_k P_node[i][k]·P_a[k][s_a]·P_b[k][s_b]; 9261 for a pitchfork
 * (cherry (a, b) and tip t below node): row (s_a*21 + s_b)*21 + s_t. States 0..19, 20 = missing data. Parity hook for the
 * transition matrices and the inner loop of JointBranchReconstruction.java:208-230. */
int sr_debug_cherry_table(sr_ctx *ctx, int32_t node, int32_t rate_class, double *out, int64_t cap_doubles, int32_t *n_rows);

#ifdef __cplusplus
}
#endif
#endif /* SUBRECON_B200_H */
