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
 * (recon/SiteResult.java:69-83) plus the print decision of SubRecon.java:138 — everything the reference keeps of a
 * column (it deliberately drops the 20x20 matrix, SiteResult.java:55). This is the result path to use whenever the
 * caller is going to build SiteResults: 112 bytes per column instead of 3 208.
 *
 * LOSSLESS BY CONSTRUCTION: the 400 probabilities sum to 1, so at most floor(1/threshold) of them can be >= threshold
 * (2 at the reference's default -threshold 0.5). A record holds `compact_cap` entries (option, default SR_COMPACT_MAX
 * = 8, i.e. every threshold >= 0.125); sr_run* REFUSE (SR_ERR_ARG) a threshold whose floor(1/threshold) exceeds the
 * capacity instead of truncating. For smaller thresholds raise the capacity — sr_set_option("compact_cap", c) with
 * c = sr_compact_cap_for(threshold), up to 400 — and step through the output with sr_compact_stride(c):
 *     record = { lnl, max_ii_prob, max_prob, n_kept, interesting }            32 bytes
 *              int16 pair[c]   (padded to a multiple of 8 bytes)
 *              double prob[c]
 * `sr_compact` below is exactly that layout for c = SR_COMPACT_MAX. */
#define SR_COMPACT_MAX 8
typedef struct {
    double  lnl;                    /* marginal site log-likelihood (SiteResult.getMarginalLnL)        */
    double  max_ii_prob;            /* largest diagonal entry   (SiteResult.getMaxIIProb)             */
    double  max_prob;               /* largest entry            (SiteResult.getMaxProb)               */
    int32_t n_kept;                 /* number of entries >= threshold                                  */
    int32_t interesting;            /* maxII <= 1-threshold && maxProb >= threshold (SubRecon.java:138) */
    int16_t pair[SR_COMPACT_MAX];   /* a*20+b of the kept entries in the reference's scan order (a-major); -1 = unused */
    double  prob[SR_COMPACT_MAX];
} sr_compact;
/* Entries a record must hold for `threshold` to be lossless: min(400, floor(1/threshold)); 400 for threshold <= 0. */
int sr_compact_cap_for(double threshold);
/* Bytes per record of capacity `cap` (1..400): 32 + round_up(2*cap, 8) + 8*cap; sizeof(sr_compact) for cap = 8. */
size_t sr_compact_stride(int cap);

/* Per-kernel device timings accumulated since the last sr_profile_reset (CUDA events on the
 * launching stream). ms_* are sums over launches. */
typedef struct {
    double  ms_pbuild;  int64_t n_pbuild;       /* K1 + K1b: matrix and clade-table builders (launches of both) */
    double  ms_prune;   int64_t n_prune;        /* K2: post-order pruning (dominant)  */
    double  ms_root;    int64_t n_root;         /* K3: root-branch joint / marginal   */
    int64_t columns_pruned;                     /* columns processed by K2 launches   */
    int64_t columns_in, columns_unique;         /* "compress_patterns": columns seen / distinct patterns computed */
    double  ms_index;   int64_t n_index;        /* K0b: clade-index pre-pass (one launch per K2 launch when the tree has clade tables) */
} sr_profile;

/* Replaces: `Executors.newFixedThreadPool(nThreads)` (SubRecon.java:109) — the worker pool becomes a GPU. */
int sr_create(sr_ctx **out, int device);
void sr_destroy(sr_ctx *ctx);
const char *sr_last_error(const sr_ctx *ctx);   /* ctx may be NULL: error of the failed sr_create */

/* The same for a whole box from ONE host process (the reference is a single JVM, SubRecon.java:93): n_dev contexts, one
 * per GPU (dev_ids NULL = devices 0..n_dev-1; n_dev <= 0 = every visible device), each driven by its own host thread
 * inside the calls below. Columns are independent (SubRecon.java:112-116), so sr_multi_run_streamed hands device g the
 * contiguous block [g*ceil(n/G), ...) of the column range and the blocks' results land in the caller's arrays in place:
 * no inter-GPU traffic at all. The setters broadcast to every context. Host buffers should come from sr_host_alloc
 * (page-locked for every device). Not thread-safe; one sr_multi per caller thread. */
typedef struct sr_multi sr_multi;
int sr_multi_create(sr_multi **out, int n_dev, const int *dev_ids);
void sr_multi_destroy(sr_multi *m);
const char *sr_multi_last_error(const sr_multi *m);            /* m may be NULL: error of the failed sr_multi_create */
int sr_multi_count(const sr_multi *m);
sr_ctx *sr_multi_ctx(sr_multi *m, int i);                      /* context i, e.g. for sr_get_profile / sr_schedule_info */
int sr_multi_set_model(sr_multi *m, const double eval[20], const double evec[400], const double ievc[400], const double pi[20]);
int sr_multi_set_rates(sr_multi *m, int K, const double *rates);
int sr_multi_set_tree(sr_multi *m, int32_t n_nodes, const int32_t *child_off, const int32_t *child_idx,
                      const double *blen, const int32_t *leaf_row, int32_t root);
int sr_multi_set_option(sr_multi *m, const char *name, int64_t value);
/* Replaces the submit/join loop (SubRecon.java:112-142) on all GPUs of the box; arguments as sr_run_streamed. */
int sr_multi_run_streamed(sr_multi *m, int32_t n_taxa, int64_t n_sites, const uint8_t *states, int64_t pitch,
                          int64_t site0, int64_t n, double *lnl, double *joint, sr_compact *compact, double threshold);

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
 * compact[n] or NULL (filled with `threshold`; records of sr_compact_stride(compact_cap) bytes — SR_ERR_ARG if that
 * capacity could truncate at this threshold, see sr_compact). Host buffers; pinned ones (sr_host_alloc) are DMA'd directly. */
int sr_run(sr_ctx *ctx, int64_t site0, int64_t n, double *lnl, double *joint, sr_compact *compact,
           double threshold);

/* Same computation with DEVICE output pointers, enqueued on `stream` (a cudaStream_t, may be NULL = the
 * context's own stream) and not synchronised: for callers that keep results in HBM. */
int sr_run_device(sr_ctx *ctx, int64_t site0, int64_t n, double *d_lnl, double *d_joint, sr_compact *d_compact,
                  double threshold, void *stream);

/* End-to-end variant: the alignment stays in HOST memory (states[taxon*pitch + site]); column chunks are
 * uploaded, computed and downloaded in a three-stage pipeline (three buffers per stage). Nothing needs to be resident. */
int sr_run_streamed(sr_ctx *ctx, int32_t n_taxa, int64_t n_sites, const uint8_t *states, int64_t pitch,
                    int64_t site0, int64_t n, double *lnl, double *joint, sr_compact *compact, double threshold);

/* Pinned host memory for the buffers above, page-locked for every device (JNI: wrap in a direct ByteBuffer). */
void *sr_host_alloc(size_t bytes);
void sr_host_free(void *p);

/* Tuning / introspection. Options: "compact_cap" (entries per compact record, 1..400, default 8; see sr_compact),
 * "strict_root" (0/1, default 0: reject, with SR_ERR_UNSUPPORTED from sr_set_tree / sr_set_rates, a tree whose root has a
 * tip as child when K > 1 — the reference throws on every column of such input, utils/Utils.java:56-70; off = return the
 * well-defined limit), "compress_patterns" (0/1: compute each distinct column pattern of a batch once and copy the result
 * to its duplicates — patterns are found by 128-bit hashes and every column is then compared byte for byte with its
 * pattern's representative, a block with a hash collision is computed the plain way; off by default; 2 = test mode with a
 * deliberately weak hash), "chunk_cols" / "stream_chunk_cols" (columns per launch / per pipeline stage),
 * "stream_taper" (0/1: short pipeline chunks at both ends of a streamed run, on by default), "input_chars" (0/1:
 * alignment bytes are residue letters, see sr_set_alignment), "cherry_tables" (0/1, default 1: tabulate every cherry's
 * contribution per state pair once per rate class and gather it instead of contracting it per column) with
 * "cherry_table_mb" (memory cap, default 4096), "cherry_pairs" (0/1, default 1: a step may gather several such clades) and
 * "cherry_pitchforks" (0/1, default 1: the same one level up, a cherry plus a tip, 9261 state triples), "k_fastest" (0/1,
 * default 0: work items ordered tile-major, i.e. the K rate classes of a column tile run concurrently and the states are read
 * from HBM once instead of K times — but every rate class's clade tables are then hot at once and miss L2 more often;
 * rate-major is 6 % faster at BASELINE config #3 and HBM is at 1 % of its roof either way), "profile"
 * (0/1: record per-kernel events), "rescale_every" (edges between power-of-two rescalings, 1..12, default 8). */
int sr_set_option(sr_ctx *ctx, const char *name, int64_t value);
int sr_get_profile(sr_ctx *ctx, sr_profile *out);     /* synchronises the context's streams */
int sr_profile_reset(sr_ctx *ctx);
/* Schedule facts for the current tree: number of pruning steps, stack depth needed, and the
 * algorithmic flops per column per rate class (820*E_int + 20*E_leaf + 1200). */
int sr_schedule_info(sr_ctx *ctx, int64_t *n_steps, int32_t *stack_depth, double *flops_per_column_rate);
/* Host-only (no GPU needed): the fused pruning schedule sr_set_tree would build. steps[16*i..] = {word, alignment row of
 * gather 0, 1, 2; for a clade gather g its second row (4+g), its third row (7+g) and its table's first row (10+g); 3 unused};
 * word bits: 0-1 MV (1 = R = P·A, 2 = R = pop ∘ P·A), 2-3 number of gathers before the MV, 4 first gather starts a new
 * partial, 5 one gather after the MV, 6 rescale, 7 push, 8/9 store as root child A/B, 10-12 gather g reads a collapsed
 * clade's table, 13-15 that clade is a pitchfork (three tips).
 * node_slot[v] = position of node v's parent edge in the matrix stream (per step: MV first, then tips), -1 if it has
 * none, -2-c if v is the top of collapsed clade c (a cherry: two tip children; or a pitchfork: a cherry and a tip — with
 * `cherry_tables` its contribution to its parent is tabulated per state combination and gathered like a tip's).
 * The arrays are validated exactly as sr_set_tree validates them (one parent per node, no cycle, ...).
 * This replaces the recursion order of downTreeMarginal (JointBranchReconstruction.java:191-247). */
int sr_schedule_dump(int32_t n_nodes, const int32_t *child_off, const int32_t *child_idx, const double *blen,
                     const int32_t *leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables, int64_t cap_steps,
                     int32_t *steps, int32_t *node_slot, int64_t *n_steps, int32_t *stack_depth);
/* Host-only: the same schedule as the two record streams the pruning kernel reads (subrecon_b200/csrc/prune.cuh):
 * crec[4*i..] for the consumer warps, prec[8*i..] for the TMA producer, clade_gathers[4*j..] = {row a, row b, row c or -1, 0}
 * for the index pre-pass; capacity cap_steps steps and 3*cap_steps clade gathers. tests/test_schedule.py interprets them on
 * the CPU against the oracle. */
int sr_schedule_records(int32_t n_nodes, const int32_t *child_off, const int32_t *child_idx, const double *blen,
                        const int32_t *leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables, int64_t cap_steps,
                        uint32_t *crec, uint32_t *prec, int32_t *clade_gathers, int64_t *n_steps, int64_t *n_clade_gathers);
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
