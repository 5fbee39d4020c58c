// subrecon_b200.cu — host side of libsubrecon_b200.so: context, pruning-schedule builder, launch
// pipeline and the C ABI declared in include/subrecon_b200.h. No torch types, no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/subrecon_b200.h"
#include "kernels.cuh"
#include <cub/cub.cuh>

using namespace srk;

static_assert(sizeof(sr_compact) == sizeof(CompactRec) && compact_stride(SR_COMPACT_MAX) == sizeof(sr_compact), "sr_compact layout");

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + (bytes >> 3);          // a little headroom so growing chunk sizes do not thrash
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct EvPair { cudaEvent_t a, b; int kind; int64_t cols; };

}  // namespace

struct sr_ctx {
    int device = 0;
    int n_sm = 0;
    int max_smem = 0;
    std::string err;
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_up[3] = {}, ev_done[3] = {}, ev_free[3] = {}, ev_sfree[3] = {};   // pipeline hand-offs, created once

    // model / rates
    bool has_model = false, has_rates = false, has_tree = false, has_aln = false;
    double eval[20], evec[400], ievc[400], pi[20];
    int K = 0;
    std::vector<double> rates;

    // tree + schedule
    int32_t n_nodes = 0, root = 0;
    std::vector<int32_t> child_off, child_idx, leaf_row;
    std::vector<double> blen;
    std::vector<int32_t> steps;         // fused schedule: STEP_INTS ints per step (see sr_schedule_dump)
    std::vector<int32_t> step_off;      // per step: offset (doubles) of its matrices in a rate class's stream
    std::vector<uint32_t> crec, prec;   // what the kernel reads: 4 words per step for the consumers, 8 for the producer (prune.cuh)
    std::vector<int32_t> clade_gathers; // per clade gather of the schedule, in order: {row a, row b, row c or -1, 0} (K0b)
    std::vector<int32_t> shape_hist;    // steps per consumer shape id (introspection)
    std::vector<int32_t> slot_kind, slot_off;   // per edge: layout kind + offset (doubles)
    std::vector<double> slot_blen;
    std::vector<int32_t> node_slot;     // node -> slot that applies its parent edge (-1: none)
    int64_t rate_stride = 0;            // doubles per rate class
    int need_depth = 0;
    int64_t e_int = 0, e_leaf = 0;
    int32_t max_leaf_row = -1;
    double root_blen = 0.0;
    bool sched_dirty = true, p_dirty = true;
    // collapsed cherries (an internal node whose two children are tips, not a root child): node, tip rows, edge lengths
    // clade tables, in the order the schedule uses them: a cherry (node, tips a b) or a pitchfork (node, cherry a b, tip t)
    std::vector<int32_t> cherry_node;            // the clade's top node
    std::vector<int32_t> cherry_kind;            // 2 = cherry (441 rows), 3 = pitchfork (9261 rows)
    std::vector<int64_t> cherry_rowoff;          // first row of the table inside one rate class's block
    std::vector<double> cherry_blen;             // [n][5] = {tip a, tip b, the node's own edge, (pitchfork:) cherry edge, third tip}
    int64_t ctab_rows = 0;                       // rows per rate class
    DevBuf d_ctab, d_cherry_blen, d_cherry_meta;
#ifdef SR_TIMELINE
    DevBuf d_timeline;
#endif

    // options
    int cherry_pairs = 1;               // let one step gather two cherries (siblings): trees with many cherry pairs
    int debug = 0;
    int k_fastest = 0;                  // work items tile-major (the K rate classes of a tile run concurrently) instead of rate-major
    int strict_root = 0;                // reject a tip as root child when K > 1 (the reference fails every column there)
    int compact_cap = SR_COMPACT_MAX;   // entries kept per compact record
    int cherry_pitchforks = 1;          // also tabulate (cherry, tip) clades: 9261 rows each
    int cherry_tables = 1;              // collapse cherries into 441-row lookup tables (see build_schedule)
    int64_t cherry_table_mb = 4096;     // memory cap for those tables; clades beyond it are pruned the ordinary way
    bool input_chars = false;           // alignment bytes are residue letters, translated on the device
    int compress = 0;                   // reduce every batch to its distinct column patterns first
    int rotate = 1;                     // named-barrier burst rotation among the warps of an SM sub-partition
    int64_t chunk_cols = 1 << 20;       // resident runs: columns per launch
    bool stream_taper = true;            // streamed runs: short chunks at both ends of the pipeline
    int64_t stream_chunk_cols = 1 << 17; // streamed runs: columns per pipeline stage (H2D | kernels | D2H overlap)
    int rescale_every = 8;
    bool profile = false;

    // device state
    DevBuf d_model, d_rates, d_crec, d_prec, d_cgathers, d_cidx, d_slotkind, d_slotoff, d_slotblen, d_pstream, d_proot, d_states, d_rootpart, d_rootexp, d_spill;
    DevBuf d_pat_keys[2], d_pat_idx[2], d_pat_h2, d_pat_head, d_pat_uid, d_pat_rep, d_pat_map, d_pat_tmp, d_pat_states, d_pat_lnl, d_pat_joint, d_pat_compact, d_pat_flag;
    int64_t last_unique = 0, last_total = 0;
    DevBuf d_chunk_states[3], d_chunk_lnl[3], d_chunk_joint[3], d_chunk_compact[3], d_peak;   // pipeline buffers (sr_run: 2, streamed: 3)
    int32_t aln_taxa = 0;
    int64_t aln_sites = 0, aln_pitch = 0;

    // pinned bounce buffers for pageable caller memory
    void* h_bounce[3] = {nullptr, nullptr, nullptr};
    size_t h_bounce_cap[3] = {0, 0, 0};
    // page-locked staging for the small per-rebuild uploads of ensure_pstream (model, rates, records, slots): copies from
    // it are truly asynchronous, so rebuilding the matrices never drains the stream; ev_stage marks the last copy out of it
    void* h_stage = nullptr;
    size_t h_stage_cap = 0;
    cudaEvent_t ev_stage = nullptr;
    bool stage_busy = false;

    std::vector<EvPair> events;
    sr_profile prof{};
};

namespace {

int fail(sr_ctx* c, int code, const char* fmt, ...) noexcept {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    try {
        if (c) c->err = buf; else g_create_error = buf;
    } catch (...) {
    }
    return code;
}

// Nothing throws or aborts across the ABI (include/subrecon_b200.h): the host side uses std::vector / std::string, so
// every exported function runs inside this fence.
template <typename F>
int guarded(sr_ctx* c, F&& f) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        return fail(c, SR_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(c, SR_ERR_ARG, "internal error: %s", e.what());
    } catch (...) {
        return fail(c, SR_ERR_ARG, "internal error");
    }
}

#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            cudaGetLastError();                                                                      \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? SR_ERR_OOM : SR_ERR_CUDA,             \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                            \
    } while (0)

// inside the pipeline loops: note the failure, leave the loop, let the common exit drain the streams
#define CUDA_STEP(ctx, rcvar, expr)                                                                  \
    {                                                                                                \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            cudaGetLastError();                                                                      \
            rcvar = fail(ctx, SR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            break;                                                                                   \
        }                                                                                            \
    }

// ------------------------------------------------------------------------------------------------
// The schedule as the kernel reads it (prune.cuh): everything a step needs is decided here, once per tree — the shape
// id that selects the specialised step body, the stack level a step pops from / pushes to (the walk is static, so the
// kernel keeps no stack pointer), the byte offset of every gather's source, the bytes the producer copies.
// ------------------------------------------------------------------------------------------------
int build_records(sr_ctx* c) {
    const int n_steps = (int)c->steps.size() / STEP_INTS;
    c->crec.assign((size_t)n_steps * 4, 0u);
    c->prec.assign((size_t)n_steps * 8, 0u);
    c->clade_gathers.clear();
    c->shape_hist.assign(8, 0);
    int sp = 0;
    for (int i = 0; i < n_steps; ++i) {
        const int32_t* st = &c->steps[(size_t)i * STEP_INTS];
        const int32_t w = st[0];
        const int mv = w & ST_MV_MASK, n_pre = (w >> ST_NPRE_SHIFT) & 3, post = (w & ST_NPOST) ? 1 : 0;
        const bool set0 = (w & ST_PRE0_SET) != 0;
        const int n_g = n_pre + post;
        uint32_t x = 0, off[3] = {0, 0, 0}, rows[3] = {0, 0, 0}, kinds = 0;
        int tix = 0;
        for (int g = 0; g < n_g; ++g) {
            if ((w >> (ST_CLADE_SHIFT + g)) & 1) {
                const bool tri = ((w >> (ST_TRI_SHIFT + g)) & 1) != 0;
                const int64_t byte_off = (int64_t)st[10 + g] * CHERRY_ROW_DOUBLES * 8;
                if (byte_off > (int64_t)UINT32_MAX - PITCHFORK_ROWS * CHERRY_ROW_DOUBLES * 8)
                    return fail(c, SR_ERR_UNSUPPORTED, "clade tables too large for 32-bit byte offsets; lower cherry_table_mb");
                x |= CR_CLADE0 << g;
                kinds |= 1u << g;
                off[g] = (uint32_t)byte_off;
                rows[g] = (uint32_t)(c->clade_gathers.size() / 4);
                const int32_t cg[4] = {st[1 + g], st[4 + g], tri ? st[7 + g] : -1, 0};
                c->clade_gathers.insert(c->clade_gathers.end(), cg, cg + 4);
            } else {
                off[g] = (uint32_t)((mv ? FRAG_BYTES : 0) + tix * TABLE_BYTES);
                rows[g] = (uint32_t)st[1 + g];
                ++tix;
            }
        }
        int shape = SH_GENERAL;
        if (n_pre == 0 && mv == 1) shape = post ? SH_MV1_POST : SH_MV1;
        else if (n_pre == 0 && mv == 2) shape = post ? SH_MV2_POST : SH_MV2;
        else if (n_pre == 2 && set0 && mv == 1) shape = post ? SH_NEW2_MV1_POST : SH_NEW2_MV1;
        else if (n_pre == 2 && set0 && mv == 2 && !post) shape = SH_NEW2_MV2;
        x |= (uint32_t)shape;
        ++c->shape_hist[shape];
        if (w & STF_RESCALE) x |= CR_RESCALE;
        if (mv == 2) {
            if (sp <= 0) return fail(c, SR_ERR_UNSUPPORTED, "internal: schedule pops an empty stack at step %d", i);
            --sp;
            x |= (uint32_t)sp << CR_POP_SHIFT;
        }
        if (w & STF_PUSH) {
            if (sp >= 31) return fail(c, SR_ERR_UNSUPPORTED, "tree needs more than 31 stacked partials");
            x |= CR_PUSH | ((uint32_t)sp << CR_PUSH_SHIFT);
            ++sp;
        }
        if (w & STF_STORE_A) x |= CR_STORE_A;
        if (w & STF_STORE_B) x |= CR_STORE_B;
        x |= (uint32_t)n_pre << CR_NPRE_SHIFT;
        if (set0) x |= CR_PRE_SET;
        x |= (uint32_t)mv << CR_MV_SHIFT;
        if (post) x |= CR_POST;
        uint32_t* cr = &c->crec[(size_t)i * 4];
        cr[0] = x; cr[1] = off[0]; cr[2] = off[1]; cr[3] = off[2];
        uint32_t* pr = &c->prec[(size_t)i * 8];
        pr[0] = (uint32_t)((int64_t)c->step_off[i] * 8);
        pr[1] = (uint32_t)((mv ? FRAG_BYTES : 0) + tix * TABLE_BYTES);
        pr[2] = (uint32_t)n_g | (kinds << 2);
        pr[3] = rows[0]; pr[4] = rows[1]; pr[5] = rows[2];
    }
    if ((int64_t)c->rate_stride * 8 > (int64_t)UINT32_MAX) return fail(c, SR_ERR_UNSUPPORTED, "tree too large for 32-bit matrix-stream offsets");
    return SR_OK;
}

// ------------------------------------------------------------------------------------------------
// Pruning schedule. The reference recurses (JointBranchReconstruction.java:191-247); here the tree is
// flattened once into a linear list of steps for a one-register-set + stack machine:
//   children of a node are visited internal-first, ordered by descending stack need (Strahler order),
//   so a binary tree of N tips needs at most ceil(log2 N) stacked partials and a caterpillar none.
// Every step applies exactly one edge (one transition matrix), so step i consumes slot i of the
// per-rate matrix stream.
// ------------------------------------------------------------------------------------------------
int build_schedule(sr_ctx* c) {
    const int n = c->n_nodes;
    const auto& off = c->child_off;
    const auto& idx = c->child_idx;
    auto is_leaf = [&](int v) { return off[v] == off[v + 1]; };
    const int rc0 = idx[off[c->root]], rc1 = idx[off[c->root] + 1];

    // post-order over the whole tree (iterative) to compute the stack need of every internal node
    std::vector<int> order;
    order.reserve(n);
    {
        std::vector<int> st{c->root};
        while (!st.empty()) {
            int v = st.back(); st.pop_back();
            order.push_back(v);
            for (int e = off[v]; e < off[v + 1]; ++e) st.push_back(idx[e]);
        }
    }
    if ((int)order.size() != n) return fail(c, SR_ERR_ARG, "tree is not connected: %d of %d nodes reachable from the root", (int)order.size(), n);
    // Cherry tables. For an internal node v with exactly two tip children a, b (and v not a root child), the vector v hands
    // to its parent, u[i] = sum_k P_v[i][k] * P_a[k][s_a] * P_b[k][s_b], depends on the column only through the state pair
    // (s_a, s_b): 21 x 21 possibilities. K1 tabulates u for all of them once per rate class (CHERRY_ROWS x 20 doubles), and
    // the pruning kernel treats v as a tip with 441 states: two gathers and one 20x20 contraction per column become one
    // gather. A Yule tree has ~N/3 cherries, a balanced one N/2, so a third to a half of all contractions disappear.
    // Pitchforks: a node whose children are a collapsed cherry (a, b) and a tip t gets the same treatment one level up —
    // 21^3 state triples, 9261 rows — and absorbs the cherry: another contraction and another gather per column gone.
    std::vector<int32_t> cherry_of(n, -1);       // 2 / 3 = collapsed cherry / pitchfork (replaced by the table index below)
    c->cherry_node.clear(); c->cherry_kind.clear(); c->cherry_rowoff.clear(); c->cherry_blen.clear();
    c->ctab_rows = 0;
    if (c->cherry_tables) {
        const int64_t row_bytes = (int64_t)std::max(1, c->K) * CHERRY_ROW_DOUBLES * (int64_t)sizeof(double);
        int64_t rows_left = (c->cherry_table_mb << 20) / row_bytes;
        int n_tab = 0;
        const int max_tab = 1 << 20;
        for (int v = 0; v < n && rows_left >= CHERRY_ROWS && n_tab < max_tab; ++v) {
            if (v == c->root || v == rc0 || v == rc1 || off[v + 1] - off[v] != 2) continue;
            if (!is_leaf(idx[off[v]]) || !is_leaf(idx[off[v] + 1])) continue;
            cherry_of[v] = 2; rows_left -= CHERRY_ROWS; ++n_tab;
        }
        if (c->cherry_pitchforks) {
            for (int v = 0; v < n && rows_left >= PITCHFORK_ROWS - CHERRY_ROWS; ++v) {
                if (v == c->root || v == rc0 || v == rc1 || off[v + 1] - off[v] != 2) continue;
                const int x = idx[off[v]], y = idx[off[v] + 1];
                const int ch = cherry_of[x] == 2 ? x : (cherry_of[y] == 2 ? y : -1);
                if (ch < 0 || !is_leaf(ch == x ? y : x)) continue;
                cherry_of[v] = 3; cherry_of[ch] = -1; rows_left -= PITCHFORK_ROWS - CHERRY_ROWS;
            }
        }
    }
    auto is_tip = [&](int v) { return is_leaf(v) || cherry_of[v] >= 0; };
    std::vector<int> need(n, 0);
    std::vector<int32_t> ord(idx.size());      // per node: children reordered (internal by need desc, then tips)
    for (int oi = n - 1; oi >= 0; --oi) {
        const int v = order[oi];
        if (is_tip(v)) continue;
        std::vector<int> inner, tips;
        for (int e = off[v]; e < off[v + 1]; ++e) (is_tip(idx[e]) ? tips : inner).push_back(idx[e]);
        std::stable_sort(inner.begin(), inner.end(), [&](int a, int b) { return need[a] > need[b]; });
        int nd = 0;
        for (size_t i = 0; i < inner.size(); ++i) nd = std::max(nd, need[inner[i]] + (i > 0 ? 1 : 0));
        need[v] = nd;
        int e = off[v];
        for (int x : inner) ord[e++] = x;
        for (int x : tips) ord[e++] = x;
    }

    // ---- primitive edge operations in execution order
    enum { P_LEAF_SET, P_LEAF_MUL, P_MV_SET, P_MV_MULPOP };
    struct Prim { int code; int child; bool identity, push, store_a, store_b; };
    std::vector<Prim> prims;
    prims.reserve(n);
    struct Frame { int v; int i; };              // i = next position in ord[] of v
    for (int which = 0; which < 2; ++which) {
        const int r = which == 0 ? rc0 : rc1;
        if (is_leaf(r)) {
            prims.push_back({P_LEAF_SET, r, true, false, false, false});
        } else {
            std::vector<Frame> fr{{r, 0}};
            while (!fr.empty()) {
                Frame& f = fr.back();
                const int v = f.v, deg = off[v + 1] - off[v];
                if (f.i == deg) {                          // node complete: hand the partial to the parent
                    fr.pop_back();
                    if (!fr.empty()) {
                        const bool first = (fr.back().i == 1);    // v was ord[0] of its parent
                        prims.push_back({first ? P_MV_SET : P_MV_MULPOP, v, false, false, false, false});
                    }
                    continue;
                }
                const int ch = ord[off[v] + f.i];
                if (!is_tip(ch)) {
                    if (f.i > 0) prims.back().push = true;        // keep the partial while the sibling subtree runs
                    ++f.i;
                    fr.push_back({ch, 0});
                } else {
                    prims.push_back({f.i == 0 ? P_LEAF_SET : P_LEAF_MUL, ch, false, false, false, false});
                    ++f.i;
                }
            }
        }
        (which == 0 ? prims.back().store_a : prims.back().store_b) = true;
    }

    // clade tables are laid out in the order the schedule uses them
    for (const Prim& pr : prims) {
        if (pr.code > P_LEAF_MUL || cherry_of[pr.child] < 0) continue;
        const int v = pr.child, kind = cherry_of[v];
        int a, b, t = -1, ce = -1;
        if (kind == 2) { a = idx[off[v]]; b = idx[off[v] + 1]; }
        else {
            const int x = idx[off[v]], y = idx[off[v] + 1];
            ce = is_leaf(x) ? y : x; t = is_leaf(x) ? x : y;
            a = idx[off[ce]]; b = idx[off[ce] + 1];
        }
        cherry_of[v] = (int32_t)c->cherry_node.size();
        c->cherry_node.push_back(v);
        c->cherry_kind.push_back(kind);
        c->cherry_rowoff.push_back(c->ctab_rows);
        c->ctab_rows += kind == 2 ? CHERRY_ROWS : PITCHFORK_ROWS;
        const double lens[5] = {c->blen[a], c->blen[b], c->blen[v], ce >= 0 ? c->blen[ce] : 0.0, t >= 0 ? c->blen[t] : 0.0};
        c->cherry_blen.insert(c->cherry_blen.end(), lens, lens + 5);
    }
    if (c->ctab_rows > INT32_MAX) return fail(c, SR_ERR_UNSUPPORTED, "clade tables too large for 32-bit row offsets; lower cherry_table_mb");
    const int max_pseudo = c->cherry_pairs ? 3 : 1;
    // ---- fuse into steps: [<=2 tip gathers] [one MV] [<=1 tip gather]; a push/store ends its step
    c->steps.clear(); c->step_off.clear();
    c->slot_kind.clear(); c->slot_off.clear(); c->slot_blen.clear();
    c->node_slot.assign(n, -1);
    c->e_int = c->e_leaf = 0;
    c->max_leaf_row = -1;
    const int T = std::max(1, c->rescale_every);
    int s_cur = 0;                               // edges multiplied into A since its last rescale
    std::vector<int> s_stack;
    int64_t off_d = 0;
    // alignment rows of a collapsed clade's tips in table-index order (a, b[, t]); returns their number
    auto clade_rows = [&](int v, int (&rws)[3]) {
        const int kind = c->cherry_kind[cherry_of[v]];
        if (kind == 2) { rws[0] = c->leaf_row[idx[off[v]]]; rws[1] = c->leaf_row[idx[off[v] + 1]]; rws[2] = 0; return 2; }
        const int x = idx[off[v]], y = idx[off[v] + 1];
        const int ce = is_leaf(x) ? y : x, t = is_leaf(x) ? x : y;
        rws[0] = c->leaf_row[idx[off[ce]]]; rws[1] = c->leaf_row[idx[off[ce] + 1]]; rws[2] = c->leaf_row[t];
        return 3;
    };
    auto add_slot = [&](const Prim& pr) {
        if (cherry_of[pr.child] >= 0) {          // no matrices in the stream: the table lives in its own buffer
            const int tid = cherry_of[pr.child];
            int rws[3];
            const int nr = clade_rows(pr.child, rws);
            c->node_slot[pr.child] = -2 - tid;
            c->e_leaf += nr; c->e_int += nr - 1;  // the algorithmic work it stands for is unchanged
            for (int t = 0; t < nr; ++t) c->max_leaf_row = std::max(c->max_leaf_row, rws[t]);
            return;
        }
        const bool leaf = pr.code <= P_LEAF_MUL;
        c->slot_kind.push_back(pr.identity ? SLOT_IDENTITY : (leaf ? SLOT_TABLE : SLOT_FRAG));
        c->slot_off.push_back((int32_t)off_d);
        c->slot_blen.push_back(c->blen[pr.child]);
        c->node_slot[pr.child] = (int)c->slot_kind.size() - 1;
        off_d += leaf ? TABLE_DOUBLES : FRAG_DOUBLES;
        if (!pr.identity) { if (leaf) ++c->e_leaf; else ++c->e_int; }
        if (leaf) c->max_leaf_row = std::max(c->max_leaf_row, c->leaf_row[pr.child]);
    };
    size_t i = 0;
    while (i < prims.size()) {
        int32_t word = 0, rows[3] = {0, 0, 0}, row_b[3] = {0, 0, 0}, row_c[3] = {0, 0, 0}, row_off[3] = {0, 0, 0};
        int n_leaf = 0, n_pre = 0;
        int n_pseudo = 0;                        // cherry-table gathers of this step (at most max_pseudo)
        auto is_pseudo = [&](const Prim& pr) { return pr.code <= P_LEAF_MUL && cherry_of[pr.child] >= 0; };
        auto take_gather = [&](const Prim& pr) {
            if (is_pseudo(pr)) {
                int rws[3];
                const int nr = clade_rows(pr.child, rws);
                rows[n_leaf] = rws[0];
                row_b[n_leaf] = rws[1];
                row_c[n_leaf] = rws[2];
                row_off[n_leaf] = (int32_t)c->cherry_rowoff[cherry_of[pr.child]];
                word |= 1 << (ST_CLADE_SHIFT + n_leaf);
                if (nr == 3) word |= 1 << (ST_TRI_SHIFT + n_leaf);
                ++n_leaf;
                ++n_pseudo;
                s_cur += nr - 1;                 // two or three edges deep
            } else {
                rows[n_leaf++] = c->leaf_row[pr.child];
            }
        };
        bool ended = false;                      // a push/store flag closes the step
        const size_t first = i;
        // the step's matrices are laid out [MV fragments][tables...] — find the MV (if any) first
        size_t j = i;
        int pre = 0;
        int seen_pseudo = 0;
        while (j < prims.size() && prims[j].code <= P_LEAF_MUL && pre < 2 && !(j > i && prims[j].code == P_LEAF_SET) &&
               !(seen_pseudo >= max_pseudo && is_pseudo(prims[j]))) {
            seen_pseudo += is_pseudo(prims[j]) ? 1 : 0;
            ++pre; ++j;
            if (prims[j - 1].push || prims[j - 1].store_a || prims[j - 1].store_b) { ended = true; break; }
        }
        size_t mv_at = (size_t)-1;
        if (!ended && j < prims.size() && prims[j].code >= P_MV_SET) mv_at = j;
        c->step_off.push_back((int32_t)off_d);
        if (mv_at != (size_t)-1) add_slot(prims[mv_at]);
        for (size_t t = i; t < i + pre; ++t) {
            const Prim& pr = prims[t];
            if (t == first && pr.code == P_LEAF_SET) { word |= ST_PRE0_SET; s_cur = 0; }
            add_slot(pr);
            take_gather(pr);
            ++n_pre; ++s_cur;
        }
        i += pre;
        const Prim* last = &prims[i - (pre ? 1 : 0)];
        if (mv_at != (size_t)-1) {
            const Prim& pr = prims[mv_at];
            word |= (pr.code == P_MV_SET) ? 1 : 2;
            if (pr.code == P_MV_MULPOP) { s_cur += s_stack.back(); s_stack.pop_back(); }
            ++s_cur;
            last = &pr;
            i = mv_at + 1;
            ended = pr.push || pr.store_a || pr.store_b;
            if (!ended && i < prims.size() && prims[i].code == P_LEAF_MUL && !(n_pseudo >= max_pseudo && is_pseudo(prims[i]))) {
                const Prim& pl = prims[i];
                add_slot(pl);
                take_gather(pl);
                word |= ST_NPOST;
                ++s_cur;
                last = &pl;
                ++i;
            }
        }
        word |= n_pre << ST_NPRE_SHIFT;
        if (s_cur >= T) { word |= STF_RESCALE; s_cur = 0; }
        if (last->push) { word |= STF_PUSH; s_stack.push_back(s_cur); }
        if (last->store_a) word |= STF_STORE_A;
        if (last->store_b) word |= STF_STORE_B;
        c->steps.push_back(word);
        c->steps.push_back(rows[0]); c->steps.push_back(rows[1]); c->steps.push_back(rows[2]);
        for (int g = 0; g < 3; ++g) c->steps.push_back(row_b[g]);
        for (int g = 0; g < 3; ++g) c->steps.push_back(row_c[g]);
        for (int g = 0; g < 3; ++g) c->steps.push_back(row_off[g]);
        for (int g = 13; g < STEP_INTS; ++g) c->steps.push_back(0);
        if (off_d > INT32_MAX - 4096) return fail(c, SR_ERR_UNSUPPORTED, "tree too large for 32-bit matrix-stream offsets");
    }
    c->rate_stride = off_d;
    c->need_depth = std::max(is_leaf(rc0) ? 0 : need[rc0], is_leaf(rc1) ? 0 : need[rc1]);
    c->root_blen = c->blen[rc0] + c->blen[rc1];
    c->sched_dirty = false;
    c->p_dirty = true;
    return build_records(c);
}

void prof_begin(sr_ctx* c, cudaStream_t s, int kind, int64_t cols) {
    if (!c->profile) return;
    EvPair e{};
    e.kind = kind; e.cols = cols;
    cudaEventCreate(&e.a); cudaEventCreate(&e.b);
    cudaEventRecord(e.a, s);
    c->events.push_back(e);
}
void prof_end(sr_ctx* c, cudaStream_t s) {
    if (!c->profile) return;
    cudaEventRecord(c->events.back().b, s);
}

// Upload schedule + model, run K1 when anything it depends on changed.
int ensure_pstream(sr_ctx* c, cudaStream_t s) {
    if (!c->has_model || !c->has_rates || !c->has_tree)
        return fail(c, SR_ERR_STATE, "model, rates and tree must be set before running");
    if (c->sched_dirty) { int rc = build_schedule(c); if (rc) return rc; }
    if (!c->p_dirty) return SR_OK;
    const int n_steps = (int)c->steps.size() / STEP_INTS, n_slots = (int)c->slot_kind.size();
    CUDA_TRY(c, c->d_model.reserve(sizeof(double) * 840));
    CUDA_TRY(c, c->d_rates.reserve(sizeof(double) * c->K));
    const int n_cg = (int)c->clade_gathers.size() / 4;
    CUDA_TRY(c, c->d_crec.reserve(sizeof(uint32_t) * 4 * n_steps));
    CUDA_TRY(c, c->d_prec.reserve(sizeof(uint32_t) * 8 * n_steps));
    CUDA_TRY(c, c->d_cgathers.reserve(sizeof(int32_t) * 4 * std::max(1, n_cg)));
    CUDA_TRY(c, c->d_slotkind.reserve(sizeof(int32_t) * n_slots));
    CUDA_TRY(c, c->d_slotoff.reserve(sizeof(int32_t) * n_slots));
    CUDA_TRY(c, c->d_slotblen.reserve(sizeof(double) * n_slots));
    CUDA_TRY(c, c->d_pstream.reserve(sizeof(double) * (size_t)c->K * c->rate_stride));
    CUDA_TRY(c, c->d_proot.reserve(sizeof(double) * (size_t)c->K * 400));
    // Every upload goes through ONE page-locked staging buffer and is enqueued on the SAME stream as the kernels that read
    // it (the context's streams are non-blocking: a plain cudaMemcpy on the legacy stream is not ordered against them, and
    // a copy from pageable memory may return before the DMA has landed). Nothing here waits for the GPU: the only host-side
    // wait is for the previous rebuild's copies to have left the staging buffer, which happened long ago.
    const int n_cherry = (int)c->cherry_node.size();
    struct Piece { void* dst; const void* src; size_t bytes; };
    double hm[840];
    memcpy(hm, c->eval, 160); memcpy(hm + 20, c->evec, 3200); memcpy(hm + 420, c->ievc, 3200); memcpy(hm + 820, c->pi, 160);
    if (n_cherry > 0) {
        CUDA_TRY(c, c->d_ctab.reserve(sizeof(double) * (size_t)c->K * c->ctab_rows * CHERRY_ROW_DOUBLES));
        CUDA_TRY(c, c->d_cherry_blen.reserve(sizeof(double) * 5 * n_cherry));
        CUDA_TRY(c, c->d_cherry_meta.reserve((sizeof(int64_t) + sizeof(int32_t)) * (size_t)n_cherry + 16));
    }
    int64_t* d_rowoff = c->d_cherry_meta.as<int64_t>();
    int32_t* d_kind = reinterpret_cast<int32_t*>(d_rowoff + n_cherry);
    const Piece pieces[] = {
        {c->d_model.p, hm, sizeof(hm)},
        {c->d_rates.p, c->rates.data(), sizeof(double) * c->K},
        {c->d_crec.p, c->crec.data(), sizeof(uint32_t) * 4 * n_steps},
        {c->d_prec.p, c->prec.data(), sizeof(uint32_t) * 8 * n_steps},
        {c->d_cgathers.p, c->clade_gathers.data(), sizeof(int32_t) * 4 * n_cg},
        {c->d_slotkind.p, c->slot_kind.data(), sizeof(int32_t) * n_slots},
        {c->d_slotoff.p, c->slot_off.data(), sizeof(int32_t) * n_slots},
        {c->d_slotblen.p, c->slot_blen.data(), sizeof(double) * n_slots},
        {c->d_cherry_blen.p, c->cherry_blen.data(), sizeof(double) * 5 * n_cherry},
        {d_rowoff, c->cherry_rowoff.data(), sizeof(int64_t) * n_cherry},
        {d_kind, c->cherry_kind.data(), sizeof(int32_t) * n_cherry},
    };
    size_t total = 0;
    for (const Piece& pc : pieces) total += (pc.bytes + 15) & ~(size_t)15;
    if (c->stage_busy) { CUDA_TRY(c, cudaEventSynchronize(c->ev_stage)); c->stage_busy = false; }
    if (total > c->h_stage_cap) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr; c->h_stage_cap = 0;
        CUDA_TRY(c, cudaHostAlloc(&c->h_stage, total + (total >> 2), cudaHostAllocPortable));
        c->h_stage_cap = total + (total >> 2);
    }
    size_t off = 0;
    for (const Piece& pc : pieces) {
        if (pc.bytes) {
            memcpy(static_cast<char*>(c->h_stage) + off, pc.src, pc.bytes);
            CUDA_TRY(c, cudaMemcpyAsync(pc.dst, static_cast<char*>(c->h_stage) + off, pc.bytes, cudaMemcpyHostToDevice, s));
        }
        off += (pc.bytes + 15) & ~(size_t)15;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_stage, s));
    c->stage_busy = true;
    PBuildParams pp{};
    pp.eval = c->d_model.as<double>();
    pp.evec = pp.eval + 20;
    pp.ievc = pp.eval + 420;
    pp.rates = c->d_rates.as<double>();
    pp.slot_kind = c->d_slotkind.as<int32_t>();
    pp.slot_off = c->d_slotoff.as<int32_t>();
    pp.slot_blen = c->d_slotblen.as<double>();
    pp.pstream = c->d_pstream.as<double>();
    pp.proot = c->d_proot.as<double>();
    pp.rate_stride = c->rate_stride;
    pp.root_blen = c->root_blen;
    pp.n_slots = n_slots;
    pp.K = c->K;
    prof_begin(c, s, 0, 0);
    pbuild_kernel<<<dim3(n_slots + 1, c->K), 128, 0, s>>>(pp);
    prof_end(c, s);
    CUDA_TRY(c, cudaGetLastError());
    if (n_cherry > 0) {
        const int64_t stride = c->ctab_rows * CHERRY_ROW_DOUBLES;
        CherryParams cp{};
        cp.eval = pp.eval; cp.evec = pp.evec; cp.ievc = pp.ievc; cp.rates = pp.rates;
        cp.blen = c->d_cherry_blen.as<double>();
        cp.kind = d_kind;
        cp.rowoff = d_rowoff;
        cp.ctab = c->d_ctab.as<double>();
        cp.rate_stride = stride;
        const int usm = CHERRY_ROWS * NS * (int)sizeof(double);              // the pitchforks' scratch (70.6 KB)
        CUDA_TRY(c, cudaFuncSetAttribute(cherry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
        prof_begin(c, s, 0, 0);
        cherry_kernel<<<dim3(n_cherry, c->K, 8), 256, usm, s>>>(cp);
        prof_end(c, s);
        CUDA_TRY(c, cudaGetLastError());
    }
    c->p_dirty = false;
    return SR_OK;
}

int launch_prune(sr_ctx* c, cudaStream_t s, PruneParams& p, int64_t cols) {
    const int fixed = P2_NSTAGE * P2_STAGE_BYTES + P2_BAR_BYTES;
    int levels = (c->max_smem - fixed) / P2_LEVEL_BYTES;
    levels = std::max(0, std::min(levels, c->need_depth));
    const int spill_levels = c->need_depth - levels;
    const int n_items = p.n_tiles * p.K;
    const int grid = std::min(n_items, c->n_sm);
    if (spill_levels > 0)
        CUDA_TRY(c, c->d_spill.reserve((size_t)grid * spill_levels * P2_LEVEL_BYTES));
    p.smem_levels = levels;
    p.spill_levels = spill_levels;
    p.spill = c->d_spill.as<double>();
    const int smem = fixed + levels * P2_LEVEL_BYTES;
    CUDA_TRY(c, cudaFuncSetAttribute(prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    prof_begin(c, s, 1, cols);
    prune_kernel<<<grid, P2_THREADS, smem, s>>>(p);
    prof_end(c, s);
    CUDA_TRY(c, cudaGetLastError());
    return SR_OK;
}

// Columns [col_begin, col_begin+n) of a device block of states -> outputs (device pointers) at out0.
int run_block_plain(sr_ctx* c, cudaStream_t s, const uint8_t* d_states, int64_t pitch, int64_t col_begin, int64_t n,
                    double* d_lnl, double* d_joint, void* d_compact, double threshold, int64_t out0) {
    if (n <= 0) return SR_OK;
    int rc = ensure_pstream(c, s);
    if (rc) return rc;
    const int S = P2_S;
    const int64_t col_start = (col_begin / 16) * 16;
    const int64_t skip = col_begin - col_start;
    const int64_t tiles = (skip + n + S - 1) / S;
    if (tiles > INT32_MAX / std::max(1, c->K)) return fail(c, SR_ERR_ARG, "too many columns in one launch; lower chunk_cols");
    const int64_t ncp = tiles * S;
    const int n_cg = (int)c->clade_gathers.size() / 4;
    CUDA_TRY(c, c->d_rootpart.reserve(sizeof(double) * 2 * (size_t)c->K * ncp * NS));
    CUDA_TRY(c, c->d_rootexp.reserve(sizeof(int32_t) * (size_t)c->K * ncp));
    CUDA_TRY(c, c->d_cidx.reserve(sizeof(uint16_t) * (size_t)std::max(1, n_cg) * ncp));
    if (n_cg) {
        // K0b: one table-row index per (clade gather, column). The rows are padded past the launch's last column (the
        // alignment block's pitch leaves >= 256 bytes of slack, filled with the missing-data code).
        const unsigned by = (unsigned)std::min<int64_t>((ncp / 2 + 255) / 256, 4096);
        prof_begin(c, s, 3, n);
        clade_index_kernel<<<dim3((unsigned)n_cg, by), 256, 0, s>>>(d_states, pitch, col_start, c->d_cgathers.as<int4>(), ncp, c->d_cidx.as<uint16_t>());
        prof_end(c, s);
        CUDA_TRY(c, cudaGetLastError());
    }
    PruneParams p{};
    p.crec = c->d_crec.as<uint4>();
    p.prec = c->d_prec.as<uint4>();
    p.pstream = c->d_pstream.as<double>();
    p.ctab = c->d_ctab.as<double>();
    p.ctab_rate_stride = c->ctab_rows * CHERRY_ROW_DOUBLES;
    p.rate_stride = c->rate_stride;
    p.states = d_states;
    p.pitch = pitch;
    p.col0 = col_start;
    p.cidx = c->d_cidx.as<uint16_t>();
    p.rootpart = c->d_rootpart.as<double>();
    p.rootexp = c->d_rootexp.as<int32_t>();
    p.ncp = ncp;
    p.n_steps = (int)c->steps.size() / STEP_INTS;
    p.n_tiles = (int)tiles;
    p.K = c->K;
    p.k_fastest = c->k_fastest;
    p.zero = 0;
#ifdef SR_TIMELINE
    if (c->d_timeline.reserve(sizeof(unsigned) * 16 * SR_TL_STEPS * 8) != cudaSuccess) return fail(c, SR_ERR_OOM, "timeline buffer");
    cudaMemsetAsync(c->d_timeline.p, 0, sizeof(unsigned) * 16 * SR_TL_STEPS * 8, s);
    p.timeline = c->d_timeline.as<unsigned>();
#endif
    rc = launch_prune(c, s, p, n);
    if (rc) return rc;
    RootParams r{};
    r.rootpart = p.rootpart;
    r.rootexp = p.rootexp;
    r.proot = c->d_proot.as<double>();
    r.pi = c->d_model.as<double>() + 820;
    r.lnl = d_lnl;
    r.joint = d_joint;
    r.compact = d_compact;
    r.compact_cap = c->compact_cap;
    r.threshold = threshold;
    r.ncp = ncp;
    r.n = n;
    r.col_skip = skip;
    r.out0 = out0;
    r.K = c->K;
    const int64_t groups = (n + ROOT_COLS - 1) / ROOT_COLS;                 // a warp takes ROOT_COLS columns at a time
    const int64_t blocks = std::min<int64_t>((groups + ROOT_WARPS - 1) / ROOT_WARPS, (int64_t)c->n_sm * 8);
    prof_begin(c, s, 2, n);
    if (c->K <= ROOT_SMEM_MAX_K) {
        const int psm = c->K * NS * NS * (int)sizeof(double);
        CUDA_TRY(c, cudaFuncSetAttribute(root_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
        root_kernel<true><<<(unsigned)blocks, ROOT_WARPS * 32, psm, s>>>(r);
    } else {
        root_kernel<false><<<(unsigned)blocks, ROOT_WARPS * 32, 0, s>>>(r);
    }
    prof_end(c, s);
    CUDA_TRY(c, cudaGetLastError());
    return SR_OK;
}


// Optional front end: hash the columns, keep one representative per distinct pattern, run the path on those and copy
// the results back to every column (bit-identical to the plain path: a column's result does not depend on its position).
int run_block(sr_ctx* c, cudaStream_t s, const uint8_t* d_states, int64_t pitch, int n_taxa, int64_t col_begin, int64_t n,
              double* d_lnl, double* d_joint, void* d_compact, double threshold, int64_t out0) {
    if (!c->compress || n < 2) return run_block_plain(c, s, d_states, pitch, col_begin, n, d_lnl, d_joint, d_compact, threshold, out0);
    if (n > INT32_MAX) return fail(c, SR_ERR_ARG, "too many columns in one launch; lower chunk_cols");
    const uint8_t* base = d_states + col_begin;
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(c, c->d_pat_keys[i].reserve(sizeof(uint64_t) * n));
        CUDA_TRY(c, c->d_pat_idx[i].reserve(sizeof(uint32_t) * n));
    }
    CUDA_TRY(c, c->d_pat_h2.reserve(sizeof(uint64_t) * n));
    CUDA_TRY(c, c->d_pat_head.reserve(sizeof(int32_t) * n));
    CUDA_TRY(c, c->d_pat_uid.reserve(sizeof(int32_t) * n));
    CUDA_TRY(c, c->d_pat_rep.reserve(sizeof(uint32_t) * n));
    CUDA_TRY(c, c->d_pat_map.reserve(sizeof(uint32_t) * n));
    const unsigned nb = (unsigned)((n + 255) / 256);
    hash_columns_kernel<<<nb, 256, 0, s>>>(base, pitch, n_taxa, n, c->d_pat_keys[0].as<uint64_t>(), c->d_pat_h2.as<uint64_t>(),
                                           c->d_pat_idx[0].as<uint32_t>(), c->compress == 2 ? std::min(n_taxa, 2) : n_taxa);
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, c->d_pat_keys[0].as<uint64_t>(), c->d_pat_keys[1].as<uint64_t>(),
                                    c->d_pat_idx[0].as<uint32_t>(), c->d_pat_idx[1].as<uint32_t>(), (int)n, 0, 64, s);
    cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, c->d_pat_head.as<int32_t>(), c->d_pat_uid.as<int32_t>(), (int)n, s);
    CUDA_TRY(c, c->d_pat_tmp.reserve(std::max(tmp_sort, tmp_scan)));
    size_t tmp_bytes = c->d_pat_tmp.cap;
    CUDA_TRY(c, cub::DeviceRadixSort::SortPairs(c->d_pat_tmp.p, tmp_bytes, c->d_pat_keys[0].as<uint64_t>(), c->d_pat_keys[1].as<uint64_t>(),
                                                c->d_pat_idx[0].as<uint32_t>(), c->d_pat_idx[1].as<uint32_t>(), (int)n, 0, 64, s));
    pattern_heads_kernel<<<nb, 256, 0, s>>>(c->d_pat_keys[1].as<uint64_t>(), c->d_pat_idx[1].as<uint32_t>(), c->d_pat_h2.as<uint64_t>(), n,
                                            c->d_pat_head.as<int32_t>());
    tmp_bytes = c->d_pat_tmp.cap;
    CUDA_TRY(c, cub::DeviceScan::InclusiveSum(c->d_pat_tmp.p, tmp_bytes, c->d_pat_head.as<int32_t>(), c->d_pat_uid.as<int32_t>(), (int)n, s));
    pattern_map_kernel<<<nb, 256, 0, s>>>(c->d_pat_idx[1].as<uint32_t>(), c->d_pat_head.as<int32_t>(), c->d_pat_uid.as<int32_t>(), n,
                                          c->d_pat_rep.as<uint32_t>(), c->d_pat_map.as<uint32_t>());
    int32_t n_uniq = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&n_uniq, c->d_pat_uid.as<int32_t>() + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));               // the launch geometry below depends on the pattern count
    c->last_unique += n_uniq;
    c->last_total += n;
    const int64_t up = (((int64_t)n_uniq + 255) / 256) * 256 + 256;
    CUDA_TRY(c, c->d_pat_states.reserve((size_t)up * n_taxa + 512));
    gather_patterns_kernel<<<dim3((unsigned)((up + 255) / 256), n_taxa), 256, 0, s>>>(base, pitch, n_taxa, c->d_pat_rep.as<uint32_t>(), n_uniq,
                                                                                   c->d_pat_states.as<uint8_t>(), up);
    // the hashes' verdict is checked against the bytes; the flag is read after the block's kernels have been enqueued
    CUDA_TRY(c, c->d_pat_flag.reserve(sizeof(int)));
    CUDA_TRY(c, cudaMemsetAsync(c->d_pat_flag.p, 0, sizeof(int), s));
    verify_patterns_kernel<<<dim3(nb, (unsigned)((n_taxa + 31) / 32)), 256, 0, s>>>(base, pitch, n_taxa, n, c->d_pat_map.as<uint32_t>(),
                                                                                  c->d_pat_states.as<uint8_t>(), up, c->d_pat_flag.as<int>());
    CUDA_TRY(c, c->d_pat_lnl.reserve(sizeof(double) * n_uniq));
    if (d_joint) CUDA_TRY(c, c->d_pat_joint.reserve(sizeof(double) * 400 * (size_t)n_uniq));
    if (d_compact) CUDA_TRY(c, c->d_pat_compact.reserve((size_t)compact_stride(c->compact_cap) * (size_t)n_uniq));
    int rc = run_block_plain(c, s, c->d_pat_states.as<uint8_t>(), up, 0, n_uniq, c->d_pat_lnl.as<double>(),
                             d_joint ? c->d_pat_joint.as<double>() : nullptr, d_compact ? c->d_pat_compact.p : nullptr, threshold, 0);
    if (rc) return rc;
    scatter_patterns_kernel<<<c->n_sm * 16, 256, 0, s>>>(c->d_pat_map.as<uint32_t>(), n, out0, c->d_pat_lnl.as<double>(),
                                                         d_joint ? c->d_pat_joint.as<double>() : nullptr,
                                                         d_compact ? c->d_pat_compact.as<uint64_t>() : nullptr, d_lnl, d_joint,
                                                         reinterpret_cast<uint64_t*>(d_compact), compact_stride(c->compact_cap) / 8);
    CUDA_TRY(c, cudaGetLastError());
    int mismatch = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&mismatch, c->d_pat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    if (mismatch) {          // two different columns shared a 128-bit hash: this block is computed column by column instead
        c->last_unique += n - n_uniq;
        return run_block_plain(c, s, d_states, pitch, col_begin, n, d_lnl, d_joint, d_compact, threshold, out0);
    }
    return SR_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

int compact_cap_for(double threshold) {
    if (!(threshold > 0.0)) return 400;
    const double k = std::floor(1.0 / threshold + 1e-9);            // entries >= threshold in a matrix that sums to 1
    return k >= 400.0 ? 400 : std::max(1, (int)k);
}

int check_run_args(sr_ctx* c, int64_t site0, int64_t n, int64_t n_sites, const void* lnl, const void* compact = nullptr, double threshold = 0.5) {
    if (!c) return SR_ERR_ARG;
    if (compact && compact_cap_for(threshold) > c->compact_cap)
        return fail(c, SR_ERR_ARG, "threshold %g can keep up to %d entries per column but a compact record holds %d: raise option "
                                   "\"compact_cap\" (records are sr_compact_stride(cap) bytes) or ask for the full joint matrix",
                    threshold, compact_cap_for(threshold), c->compact_cap);
    if (site0 < 0 || n < 0 || site0 + n > n_sites)
        return fail(c, SR_ERR_ARG, "column range [%lld, %lld) outside the alignment (%lld sites)", (long long)site0,
                    (long long)(site0 + n), (long long)n_sites);
    if (n > 0 && !lnl) return fail(c, SR_ERR_ARG, "lnl output is required");
    return SR_OK;
}

}  // namespace

// ================================================================================================ ABI
namespace {


static int sr_create_impl(sr_ctx** out, int device) {
    if (!out) return fail(nullptr, SR_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, SR_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n_dev) return fail(nullptr, SR_ERR_ARG, "device %d out of range (0..%d)", device, n_dev - 1);
    sr_ctx* c = new (std::nothrow) sr_ctx();
    if (!c) return fail(nullptr, SR_ERR_OOM, "out of host memory");
    c->device = device;
    cudaDeviceProp pr{};
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&pr, device)) != cudaSuccess) {
        delete c;
        return fail(nullptr, SR_ERR_CUDA, "cannot open device %d: %s", device, cudaGetErrorString(e));
    }
    if (pr.major != 10) {
        delete c;
        return fail(nullptr, SR_ERR_UNSUPPORTED, "device %d is sm_%d%d; this build targets sm_100a (B200) only", device, pr.major, pr.minor);
    }
    c->n_sm = pr.multiProcessorCount;
    c->max_smem = (int)pr.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return fail(nullptr, SR_ERR_CUDA, "cannot create CUDA streams");
    }
    for (int i = 0; i < 3; ++i) {
        if (cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_sfree[i], cudaEventDisableTiming) != cudaSuccess) {
            sr_destroy(c);
            return fail(nullptr, SR_ERR_CUDA, "cannot create CUDA events");
        }
    }
    if (cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming) != cudaSuccess) {
        sr_destroy(c);
        return fail(nullptr, SR_ERR_CUDA, "cannot create CUDA events");
    }
    *out = c;
    return SR_OK;
}

static int sr_set_model_impl(sr_ctx* c, const double* eval, const double* evec, const double* ievc, const double* pi) {
    if (!c) return SR_ERR_ARG;
    if (!eval || !evec || !ievc || !pi) return fail(c, SR_ERR_ARG, "NULL model array");
    for (int i = 0; i < 20; ++i)
        if (!(pi[i] >= 0.0) || !std::isfinite(eval[i])) return fail(c, SR_ERR_ARG, "pi/eval entry %d is not a finite non-negative number", i);
    memcpy(c->eval, eval, 160); memcpy(c->evec, evec, 3200); memcpy(c->ievc, ievc, 3200); memcpy(c->pi, pi, 160);
    c->has_model = true;
    c->p_dirty = true;
    return SR_OK;
}

static int check_strict_root(sr_ctx* c);

static int sr_set_rates_impl(sr_ctx* c, int K, const double* rates) {
    if (!c) return SR_ERR_ARG;
    if (K < 1) return fail(c, SR_ERR_ARG, "ERROR: -n (number of rate categories) must be 1 or higher");   // SubRecon.java:209-210
    if (!rates) return fail(c, SR_ERR_ARG, "rates is NULL");
    for (int i = 0; i < K; ++i)
        if (!std::isfinite(rates[i]) || rates[i] < 0.0) return fail(c, SR_ERR_ARG, "rate %d is not a finite non-negative number", i);
    if (K != c->K) c->sched_dirty = true;            // the cherry-table memory cap is per rate class
    const int old_K = c->K;
    c->K = K;
    c->has_rates = true;
    if (int rc = check_strict_root(c)) { c->K = old_K; c->has_rates = !c->rates.empty(); return rc; }
    c->rates.assign(rates, rates + K);
    c->p_dirty = true;
    return SR_OK;
}

// Everything sr_set_tree and the host-only schedule exports accept: CSR arrays that describe ONE rooted tree. Every
// non-root node must be the child of exactly one parent and be reachable from the root (so no cycle, no node shared
// by two parents, no orphan), no node has a single child (PAL: "Node with single child enountered"), the root has
// exactly two children (SubRecon.java:212-213), every tip has an alignment row. `c` only receives the message.
static int validate_tree(sr_ctx* c, int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                         const int32_t* leaf_row, int32_t root) {
    if (n_nodes < 3 || !child_off || !child_idx || !blen || !leaf_row) return fail(c, SR_ERR_ARG, "bad tree arrays");
    if (root < 0 || root >= n_nodes) return fail(c, SR_ERR_ARG, "root index out of range");
    if (child_off[0] != 0) return fail(c, SR_ERR_ARG, "child_off[0] must be 0");
    for (int v = 0; v < n_nodes; ++v) {
        const int64_t deg = (int64_t)child_off[v + 1] - child_off[v];
        if (deg < 0) return fail(c, SR_ERR_ARG, "child_off is not monotone at node %d", v);
        if (deg == 1) return fail(c, SR_ERR_ARG, "Node with single child enountered (node %d)", v);   // PAL ReadTree::readNH
        if (deg == 0 && leaf_row[v] < 0) return fail(c, SR_ERR_ARG, "tip node %d has no alignment row (taxon missing from the alignment)", v);
        if (deg == 0 && leaf_row[v] >= (1 << 23)) return fail(c, SR_ERR_UNSUPPORTED, "alignment row %d exceeds 2^23", leaf_row[v]);
        if (!(blen[v] >= 0.0) && v != root) return fail(c, SR_ERR_ARG, "branch length of node %d is negative or NaN", v);
    }
    const int n_edges = child_off[n_nodes];
    if (n_edges != n_nodes - 1) return fail(c, SR_ERR_ARG, "tree has %d edges for %d nodes", n_edges, n_nodes);
    std::vector<uint8_t> n_parents(n_nodes, 0);
    for (int e = 0; e < n_edges; ++e) {
        const int ch = child_idx[e];
        if (ch < 0 || ch >= n_nodes || ch == root) return fail(c, SR_ERR_ARG, "child_idx[%d] invalid", e);
        if (n_parents[ch]++) return fail(c, SR_ERR_ARG, "node %d is listed as a child twice: not a tree", ch);
    }
    if (child_off[root + 1] - child_off[root] != 2)
        return fail(c, SR_ERR_ARG, "ERROR: Tree root has more than two descendents. Is the tree rooted correctly?");   // SubRecon.java:212-213
    // n-1 edges and one parent per non-root node: the graph is a tree iff everything hangs off the root
    std::vector<uint8_t> seen(n_nodes, 0);
    std::vector<int32_t> st{root};
    int reached = 0;
    seen[root] = 1;
    while (!st.empty()) {
        const int v = st.back(); st.pop_back();
        ++reached;
        for (int e = child_off[v]; e < child_off[v + 1]; ++e)
            if (!seen[child_idx[e]]) { seen[child_idx[e]] = 1; st.push_back(child_idx[e]); }
    }
    if (reached != n_nodes) return fail(c, SR_ERR_ARG, "tree is not connected: %d of %d nodes reachable from the root (cycle?)", reached, n_nodes);
    return SR_OK;
}

// The reference cannot handle a tip as a root child when K > 1: every column throws inside getLnSumComponents
// (utils/Utils.java:56-70) and is skipped (SubRecon.java:127-135). With option "strict_root" that input is
// rejected up front, as SURVEY §8(b) asks; without it the native path returns the well-defined limit.
static int check_strict_root(sr_ctx* c) {
    if (!c->strict_root || !c->has_tree || !c->has_rates || c->K <= 1) return SR_OK;
    for (int e = c->child_off[c->root]; e < c->child_off[c->root + 1]; ++e) {
        const int v = c->child_idx[e];
        if (c->child_off[v] == c->child_off[v + 1])
            return fail(c, SR_ERR_UNSUPPORTED, "a child of the root is a tip (node %d) and K = %d > 1: the reference fails on every column "
                                               "of such a tree (utils/Utils.java:56-70); re-root the tree or use one rate class", v, c->K);
    }
    return SR_OK;
}

static int sr_set_tree_impl(sr_ctx* c, int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                const int32_t* leaf_row, int32_t root) {
    if (!c) return SR_ERR_ARG;
    int rc = validate_tree(c, n_nodes, child_off, child_idx, blen, leaf_row, root);
    if (rc) return rc;
    c->n_nodes = n_nodes;
    c->root = root;
    c->child_off.assign(child_off, child_off + n_nodes + 1);
    c->child_idx.assign(child_idx, child_idx + child_off[n_nodes]);
    c->blen.assign(blen, blen + n_nodes);
    c->leaf_row.assign(leaf_row, leaf_row + n_nodes);
    c->has_tree = true;
    c->sched_dirty = true;
    if ((rc = check_strict_root(c))) { c->has_tree = false; return rc; }
    return build_schedule(c);
}

static int sr_set_alignment_impl(sr_ctx* c, int32_t n_taxa, int64_t n_sites, const uint8_t* states, int64_t pitch) {
    if (!c) return SR_ERR_ARG;
    if (n_taxa < 1 || n_sites < 0 || !states || pitch < n_sites) return fail(c, SR_ERR_ARG, "bad alignment arguments");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int64_t dp = ((n_sites + 255) / 256) * 256 + 256;       // row pitch: 16-byte aligned rows + slack for the last tile
    CUDA_TRY(c, c->d_states.reserve((size_t)dp * n_taxa + 512));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_states.p, SR_STATE_MISSING, (size_t)dp * n_taxa + 512, c->stream));
    if (n_sites > 0)
        CUDA_TRY(c, cudaMemcpy2DAsync(c->d_states.p, dp, states, pitch, n_sites, n_taxa, cudaMemcpyHostToDevice, c->stream));
    if (c->input_chars) encode_chars_kernel<<<c->n_sm * 8, 256, 0, c->stream>>>(c->d_states.as<uint4>(), ((size_t)dp * n_taxa + 512) / 16);
    else sanitize_kernel<<<c->n_sm * 8, 256, 0, c->stream>>>(c->d_states.as<uint4>(), ((size_t)dp * n_taxa + 512) / 16);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->aln_taxa = n_taxa;
    c->aln_sites = n_sites;
    c->aln_pitch = dp;
    c->has_aln = true;
    return SR_OK;
}

static int sr_run_device_impl(sr_ctx* c, int64_t site0, int64_t n, double* d_lnl, double* d_joint, sr_compact* d_compact,
                  double threshold, void* stream) {
    if (!c) return SR_ERR_ARG;
    if (!c->has_aln) return fail(c, SR_ERR_STATE, "alignment not set");
    int rc = check_run_args(c, site0, n, c->aln_sites, d_lnl ? (void*)d_lnl : (void*)d_joint, d_compact, threshold);
    if (rc) return rc;
    if (c->has_tree && c->max_leaf_row >= c->aln_taxa) return fail(c, SR_ERR_ARG, "tree refers to alignment row %d but the alignment has %d taxa", c->max_leaf_row, c->aln_taxa);
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    for (int64_t done = 0; done < n;) {
        const int64_t m = std::min(n - done, c->chunk_cols);
        rc = run_block(c, s, c->d_states.as<uint8_t>(), c->aln_pitch, c->aln_taxa, site0 + done, m, d_lnl, d_joint, d_compact, threshold, done);
        if (rc) return rc;
        done += m;
    }
    return SR_OK;
}

// Results back to the caller. Page-locked destinations (sr_host_alloc) are DMA'd asynchronously on the D2H stream, which
// already waits for `done`. Pageable destinations are copied synchronously after the host itself has waited for `done`:
// asynchronous copies involving pageable memory are staged by the driver and are not something to order by stream events.
int d2h(sr_ctx* c, cudaStream_t s, cudaEvent_t done, bool pinned, void* dst, const void* src, size_t bytes) {
    if (pinned) {
        CUDA_TRY(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    } else {
        CUDA_TRY(c, cudaEventSynchronize(done));
        CUDA_TRY(c, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    }
    return SR_OK;
}

static int sr_run_impl(sr_ctx* c, int64_t site0, int64_t n, double* lnl, double* joint, sr_compact* compact, double threshold) {
    if (!c) return SR_ERR_ARG;
    if (!c->has_aln) return fail(c, SR_ERR_STATE, "alignment not set");
    int rc = check_run_args(c, site0, n, c->aln_sites, lnl, compact, threshold);
    const size_t cstride = (size_t)compact_stride(c->compact_cap);
    if (rc) return rc;
    if (c->has_tree && c->max_leaf_row >= c->aln_taxa) return fail(c, SR_ERR_ARG, "tree refers to alignment row %d but the alignment has %d taxa", c->max_leaf_row, c->aln_taxa);
    CUDA_TRY(c, cudaSetDevice(c->device));
    // chunked: compute chunk i on `stream` while chunk i-1 drains to the host on `s_d2h`
    const int64_t chunk = std::min<int64_t>(c->chunk_cols, std::max<int64_t>(n, 1));
    cudaEvent_t* done_ev = c->ev_done;
    cudaEvent_t* free_ev = c->ev_free;
    const bool pin_lnl = is_pinned(lnl), pin_joint = joint && is_pinned(joint), pin_compact = compact && is_pinned(compact);
    int ci = 0;
    rc = SR_OK;
    for (int64_t done = 0; done < n && rc == SR_OK; ++ci) {
        const int b = ci & 1;
        const int64_t m = std::min(n - done, chunk);
        if ((rc = (c->d_chunk_lnl[b].reserve(sizeof(double) * m) == cudaSuccess ? SR_OK : fail(c, SR_ERR_OOM, "device OOM (lnl chunk)")))) break;
        if (joint && c->d_chunk_joint[b].reserve(sizeof(double) * 400 * m) != cudaSuccess) { rc = fail(c, SR_ERR_OOM, "device OOM (joint chunk)"); break; }
        if (compact && c->d_chunk_compact[b].reserve(cstride * m) != cudaSuccess) { rc = fail(c, SR_ERR_OOM, "device OOM (compact chunk)"); break; }
        if (ci >= 2) CUDA_STEP(c, rc, cudaStreamWaitEvent(c->stream, free_ev[b], 0));      // buffer b drained?
        rc = run_block(c, c->stream, c->d_states.as<uint8_t>(), c->aln_pitch, c->aln_taxa, site0 + done, m, c->d_chunk_lnl[b].as<double>(),
                       joint ? c->d_chunk_joint[b].as<double>() : nullptr, compact ? c->d_chunk_compact[b].p : nullptr, threshold, 0);
        if (rc) break;
        CUDA_STEP(c, rc, cudaEventRecord(done_ev[b], c->stream));
        CUDA_STEP(c, rc, cudaStreamWaitEvent(c->s_d2h, done_ev[b], 0));
        if ((rc = d2h(c, c->s_d2h, done_ev[b], pin_lnl, lnl + done, c->d_chunk_lnl[b].p, sizeof(double) * m))) break;
        if (joint && (rc = d2h(c, c->s_d2h, done_ev[b], pin_joint, joint + done * 400, c->d_chunk_joint[b].p, sizeof(double) * 400 * m))) break;
        if (compact && (rc = d2h(c, c->s_d2h, done_ev[b], pin_compact, reinterpret_cast<char*>(compact) + cstride * done, c->d_chunk_compact[b].p, cstride * m))) break;
        CUDA_STEP(c, rc, cudaEventRecord(free_ev[b], c->s_d2h));
        done += m;
    }
    cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->s_d2h);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        return fail(c, SR_ERR_CUDA, "kernel execution failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    }
    return SR_OK;
}

static int sr_run_streamed_impl(sr_ctx* c, int32_t n_taxa, int64_t n_sites, const uint8_t* states, int64_t pitch, int64_t site0,
                    int64_t n, double* lnl, double* joint, sr_compact* compact, double threshold) {
    if (!c) return SR_ERR_ARG;
    if (n_taxa < 1 || !states || pitch < n_sites) return fail(c, SR_ERR_ARG, "bad alignment arguments");
    int rc = check_run_args(c, site0, n, n_sites, lnl, compact, threshold);
    const size_t cstride = (size_t)compact_stride(c->compact_cap);
    if (rc) return rc;
    if (c->has_tree && c->max_leaf_row >= n_taxa) return fail(c, SR_ERR_ARG, "tree refers to alignment row %d but the alignment has %d taxa", c->max_leaf_row, n_taxa);
    CUDA_TRY(c, cudaSetDevice(c->device));
    rc = ensure_pstream(c, c->stream);
    if (rc) return rc;
    // three-stage pipeline over column chunks: H2D (s_h2d) -> kernels (stream) -> D2H (s_d2h), NB buffers each. Three, not
    // two: with two, the kernels of chunk i wait for the download of chunk i-2, and when chunk i-1 is a short one (the
    // ramp at the end, the odd remainder) its kernels finish long before that download does — a whole chunk's D2H was exposed.
    constexpr int NB = 3;
    // Pipeline chunk: about stream_chunk_cols columns, rounded to a whole number of persistent-kernel waves
    // (tiles x K work items must be a multiple of the SM count, or the last wave of every chunk runs part empty).
    int64_t chunk;
    {
        const int64_t S = P2_S;
        int64_t g = c->n_sm, b = std::max(1, c->K);
        while (b) { const int64_t t = g % b; g = b; b = t; }
        const int64_t unit = (c->n_sm / g) * S;                       // columns per whole wave set
        const int64_t target = std::min<int64_t>(std::min(c->chunk_cols, c->stream_chunk_cols), std::max<int64_t>(n, 1));
        chunk = (target >= unit) ? (target / unit) * unit : (target + 255) / 256 * 256;
        chunk = std::max<int64_t>(256, (chunk + 15) / 16 * 16);
    }
    const int64_t cpitch = (chunk + 255) / 256 * 256 + 256;
    cudaEvent_t* up_ev = c->ev_up;
    cudaEvent_t* done_ev = c->ev_done;
    cudaEvent_t* free_ev = c->ev_free;
    cudaEvent_t* sfree_ev = c->ev_sfree;
    const bool src_pinned = is_pinned(states);
    const bool pin_lnl = is_pinned(lnl), pin_joint = joint && is_pinned(joint), pin_compact = compact && is_pinned(compact);
    // Chunk sizes: full chunks in the middle, a ramp of short ones at both ends. The pipeline's exposed time is the
    // upload of the first chunk plus the download of the last one (3200 B per column of joint output).
    std::vector<int64_t> sizes;
    {
        const int64_t S = P2_S;
        int64_t g = c->n_sm, b = std::max(1, c->K);
        while (b) { const int64_t t = g % b; g = b; b = t; }
        const int64_t unit = (c->n_sm / g) * S;
        // head: 1/8, 1/4, 1/2 of a chunk (the first upload is what is exposed). tail: a geometric ramp with ratio 3/4 down to
        // 1/8 — a download takes about two thirds of the time its chunk's kernels take, so each chunk's download is over
        // before the next, 3/4 as long, has been computed, and only the last small one is exposed.
        std::vector<int64_t> head, tail;
        int64_t head_sum = 0, tail_sum = 0;
        if (c->stream_taper && chunk >= 8 * unit) {
            for (int i = 3; i >= 1; --i) { head.push_back(std::max<int64_t>(unit, (chunk >> i) / unit * unit)); head_sum += head.back(); }
            for (int64_t t = chunk * 3 / 4; t >= chunk / 8; t = t * 3 / 4) { tail.push_back(std::max<int64_t>(unit, t / unit * unit)); tail_sum += tail.back(); }
        }
        int64_t left = n;
        if (!head.empty() && n >= head_sum + tail_sum + chunk) {
            for (int64_t h : head) { sizes.push_back(h); left -= h; }
            const int64_t odd = (left - tail_sum) % chunk;                 // the remainder goes in front of the full chunks
            if (odd > 0) { sizes.push_back(odd); left -= odd; }
            while (left - tail_sum >= chunk) { sizes.push_back(chunk); left -= chunk; }
            for (int64_t t : tail) { sizes.push_back(t); left -= t; }
        } else {
            while (left > 0) { const int64_t m = std::min(left, chunk); sizes.push_back(m); left -= m; }
        }
    }
    const int64_t mmax = std::min(n, chunk);
    for (int b = 0; b < NB && rc == SR_OK; ++b) {
        if (c->d_chunk_states[b].reserve((size_t)cpitch * n_taxa + 512) != cudaSuccess ||
            c->d_chunk_lnl[b].reserve(sizeof(double) * mmax) != cudaSuccess ||
            (joint && c->d_chunk_joint[b].reserve(sizeof(double) * 400 * mmax) != cudaSuccess) ||
            (compact && c->d_chunk_compact[b].reserve(cstride * mmax) != cudaSuccess)) {
            cudaGetLastError();
            rc = fail(c, SR_ERR_OOM, "device OOM (pipeline chunk buffers)");
        }
        if ((int)sizes.size() <= b + 1) break;
    }
#ifdef SR_TIMELINE
    struct TlEv { cudaEvent_t h0, h1, k0, k1, d0, d1; int64_t m; };
    std::vector<TlEv> tl(sizes.size());
    for (auto& e : tl) { cudaEventCreate(&e.h0); cudaEventCreate(&e.h1); cudaEventCreate(&e.k0); cudaEventCreate(&e.k1); cudaEventCreate(&e.d0); cudaEventCreate(&e.d1); }
#define SR_PIPE_EV(ev, st) cudaEventRecord(tl[ci].ev, st)
    auto tl_set_m = [&](size_t i, int64_t mm) { tl[i].m = mm; };
#else
#define SR_PIPE_EV(ev, st) do { } while (0)
    auto tl_set_m = [](size_t, int64_t) {};
#endif
    int64_t done = 0;
    for (size_t ci = 0; ci < sizes.size() && rc == SR_OK; ++ci) {
        const int b = (int)(ci % NB);
        const int64_t m = sizes[ci];
        if (ci >= NB) CUDA_STEP(c, rc, cudaStreamWaitEvent(c->s_h2d, sfree_ev[b], 0));    // kernels of chunk ci-NB have consumed states buffer b
        const uint8_t* src = states + site0 + done;
        int64_t src_pitch = pitch;
        if (!src_pinned) {
            // Pageable caller memory is never handed to an asynchronous copy (its ordering against the other streams'
            // work is the driver's business, and it showed: rare stale state bytes at 4M+ columns). Stage the chunk in a
            // page-locked bounce buffer on the host thread instead; the DMA then runs from pinned memory like any other.
            const size_t need = (size_t)mmax * n_taxa;                   // sized for a full chunk once, not per ramp step
            if (c->h_bounce_cap[b] < need) {
                if (c->h_bounce[b]) {
                    if (ci >= NB) CUDA_STEP(c, rc, cudaEventSynchronize(up_ev[b]));
                    cudaFreeHost(c->h_bounce[b]); c->h_bounce[b] = nullptr; c->h_bounce_cap[b] = 0;
                }
                if (cudaMallocHost(&c->h_bounce[b], need) != cudaSuccess) { cudaGetLastError(); rc = fail(c, SR_ERR_OOM, "cannot allocate the pinned staging buffer"); break; }
                c->h_bounce_cap[b] = need;
            } else if (ci >= NB) {
                CUDA_STEP(c, rc, cudaEventSynchronize(up_ev[b]));        // the DMA out of this bounce buffer (chunk ci-NB) is done
            }
            uint8_t* hb = static_cast<uint8_t*>(c->h_bounce[b]);
            for (int t = 0; t < n_taxa; ++t) memcpy(hb + (size_t)t * m, src + (size_t)t * pitch, (size_t)m);
            src = hb;
            src_pitch = m;
        }
        SR_PIPE_EV(h0, c->s_h2d);
        cudaError_t e = cudaMemcpy2DAsync(c->d_chunk_states[b].p, cpitch, src, src_pitch, m, n_taxa, cudaMemcpyHostToDevice, c->s_h2d);
        if (e != cudaSuccess) { rc = fail(c, SR_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e)); break; }
        if (c->input_chars) encode_chars_kernel<<<c->n_sm * 4, 256, 0, c->s_h2d>>>(c->d_chunk_states[b].as<uint4>(), ((size_t)cpitch * n_taxa + 512) / 16);
        else sanitize_kernel<<<c->n_sm * 4, 256, 0, c->s_h2d>>>(c->d_chunk_states[b].as<uint4>(), ((size_t)cpitch * n_taxa + 512) / 16);
        CUDA_STEP(c, rc, cudaGetLastError());
        CUDA_STEP(c, rc, cudaEventRecord(up_ev[b], c->s_h2d));
        SR_PIPE_EV(h1, c->s_h2d);
        CUDA_STEP(c, rc, cudaStreamWaitEvent(c->stream, up_ev[b], 0));
        if (ci >= NB) CUDA_STEP(c, rc, cudaStreamWaitEvent(c->stream, free_ev[b], 0));    // output buffers b drained
        SR_PIPE_EV(k0, c->stream);
        rc = run_block(c, c->stream, c->d_chunk_states[b].as<uint8_t>(), cpitch, n_taxa, 0, m, c->d_chunk_lnl[b].as<double>(),
                       joint ? c->d_chunk_joint[b].as<double>() : nullptr, compact ? c->d_chunk_compact[b].p : nullptr, threshold, 0);
        if (rc) break;
        CUDA_STEP(c, rc, cudaEventRecord(done_ev[b], c->stream));
        CUDA_STEP(c, rc, cudaEventRecord(sfree_ev[b], c->stream));
        SR_PIPE_EV(k1, c->stream);
        CUDA_STEP(c, rc, cudaStreamWaitEvent(c->s_d2h, done_ev[b], 0));
        SR_PIPE_EV(d0, c->s_d2h);
        if ((rc = d2h(c, c->s_d2h, done_ev[b], pin_lnl, lnl + done, c->d_chunk_lnl[b].p, sizeof(double) * m))) break;
        if (joint && (rc = d2h(c, c->s_d2h, done_ev[b], pin_joint, joint + done * 400, c->d_chunk_joint[b].p, sizeof(double) * 400 * m))) break;
        if (compact && (rc = d2h(c, c->s_d2h, done_ev[b], pin_compact, reinterpret_cast<char*>(compact) + cstride * done, c->d_chunk_compact[b].p, cstride * m))) break;
        CUDA_STEP(c, rc, cudaEventRecord(free_ev[b], c->s_d2h));
        SR_PIPE_EV(d1, c->s_d2h);
        tl_set_m(ci, m);
        done += m;
    }
    cudaError_t e0 = cudaStreamSynchronize(c->s_h2d), e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->s_d2h);
#ifdef SR_TIMELINE
    if (getenv("SR_PIPE_DUMP")) {
        for (size_t i = 0; i < tl.size(); ++i) {
            float t[6];
            cudaEvent_t evs[6] = {tl[i].h0, tl[i].h1, tl[i].k0, tl[i].k1, tl[i].d0, tl[i].d1};
            for (int j = 0; j < 6; ++j) cudaEventElapsedTime(&t[j], tl[0].h0, evs[j]);
            fprintf(stderr, "chunk %2zu cols %7lld | h2d %7.2f-%7.2f | kernels %7.2f-%7.2f | d2h %7.2f-%7.2f\n", i, (long long)tl[i].m, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
    }
    for (auto& e : tl) { cudaEventDestroy(e.h0); cudaEventDestroy(e.h1); cudaEventDestroy(e.k0); cudaEventDestroy(e.k1); cudaEventDestroy(e.d0); cudaEventDestroy(e.d1); }
#endif
    if (rc) return rc;
    if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        return fail(c, SR_ERR_CUDA, "pipeline execution failed: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2)));
    }
    return SR_OK;
}



static int sr_set_option_impl(sr_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return SR_ERR_ARG;
    const std::string k(name);
    if (k == "debug") {
        c->debug = (int)value;
    } else if (k == "compact_cap") {
        if (value < 1 || value > 400) return fail(c, SR_ERR_ARG, "compact_cap must be in 1..400");
        c->compact_cap = (int)value;
    } else if (k == "k_fastest") {
        c->k_fastest = value != 0;
    } else if (k == "strict_root") {
        c->strict_root = value != 0;
    } else if (k == "chunk_cols") {
        if (value < 256) return fail(c, SR_ERR_ARG, "chunk_cols must be >= 256");
        c->chunk_cols = (value + 255) / 256 * 256;
    } else if (k == "compress_patterns") {
        if (value < 0 || value > 2) return fail(c, SR_ERR_ARG, "compress_patterns must be 0, 1 or 2");
        c->compress = (int)value;                      // 2 = test mode: hash only the first two taxa, so that hashes collide
        c->last_unique = c->last_total = 0;
    } else if (k == "cherry_tables") {
        c->cherry_tables = value != 0;
        c->sched_dirty = true;
    } else if (k == "cherry_pitchforks") {
        c->cherry_pitchforks = value != 0;
        c->sched_dirty = true;
    } else if (k == "cherry_pairs") {
        c->cherry_pairs = value != 0;
        c->sched_dirty = true;
    } else if (k == "cherry_table_mb") {
        if (value < 0) return fail(c, SR_ERR_ARG, "cherry_table_mb must be >= 0");
        c->cherry_table_mb = value;
        c->sched_dirty = true;
    } else if (k == "input_chars") {
        c->input_chars = value != 0;                  // applies to the next sr_set_alignment / sr_run_streamed
    } else if (k == "stream_chunk_cols") {
        if (value < 256) return fail(c, SR_ERR_ARG, "stream_chunk_cols must be >= 256");
        c->stream_chunk_cols = (value + 255) / 256 * 256;
    } else if (k == "stream_taper") {
        c->stream_taper = value != 0;
    } else if (k == "profile") {
        c->profile = value != 0;
    } else if (k == "rescale_every") {
        // Upper bound 12: between two rescalings a partial can collect about twice the period (a popped sibling brings
        // its own un-rescaled edges, a clade gather counts two or three), each edge costing up to ~1e-13 at the 1e-9
        // branch-length clamp; 12 still passes tests/test_gpu_parity.py::test_adversarial_underflow, 16 does not.
        if (value < 1 || value > 12) return fail(c, SR_ERR_ARG, "rescale_every must be in 1..12");
        c->rescale_every = (int)value;
        c->sched_dirty = true;
    } else {
        return fail(c, SR_ERR_ARG, "unknown option '%s'", name);
    }
    return SR_OK;
}

static int sr_get_profile_impl(sr_ctx* c, sr_profile* out) {
    if (!c || !out) return SR_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaDeviceSynchronize());
    for (auto& e : c->events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            if (e.kind == 0) { c->prof.ms_pbuild += ms; ++c->prof.n_pbuild; }
            else if (e.kind == 1) { c->prof.ms_prune += ms; ++c->prof.n_prune; c->prof.columns_pruned += e.cols; }
            else if (e.kind == 3) { c->prof.ms_index += ms; ++c->prof.n_index; }
            else { c->prof.ms_root += ms; ++c->prof.n_root; }
        } else {
            cudaGetLastError();
        }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    c->events.clear();
    c->prof.columns_in = c->last_total;
    c->prof.columns_unique = c->last_unique;
    *out = c->prof;
    return SR_OK;
}

static int sr_profile_reset_impl(sr_ctx* c) {
    if (!c) return SR_ERR_ARG;
    sr_profile tmp;
    int rc = sr_get_profile(c, &tmp);
    c->prof = sr_profile{};
    c->last_unique = c->last_total = 0;
    return rc;
}

static int sr_schedule_info_impl(sr_ctx* c, int64_t* n_steps, int32_t* stack_depth, double* flops) {
    if (!c) return SR_ERR_ARG;
    if (!c->has_tree) return fail(c, SR_ERR_STATE, "tree not set");
    if (c->sched_dirty) { int rc = build_schedule(c); if (rc) return rc; }
    if (n_steps) *n_steps = (int64_t)c->steps.size() / STEP_INTS;
    if (stack_depth) *stack_depth = c->need_depth;
    if (flops) *flops = 820.0 * (double)c->e_int + 20.0 * (double)c->e_leaf + 1200.0;
    return SR_OK;
}

static int host_schedule(sr_ctx& tmp, int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                         const int32_t* leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables) {
    int rc = validate_tree(&tmp, n_nodes, child_off, child_idx, blen, leaf_row, root);
    if (rc) return rc;
    tmp.n_nodes = n_nodes;
    tmp.root = root;
    tmp.child_off.assign(child_off, child_off + n_nodes + 1);
    tmp.child_idx.assign(child_idx, child_idx + child_off[n_nodes]);
    tmp.blen.assign(blen, blen + n_nodes);
    tmp.leaf_row.assign(leaf_row, leaf_row + n_nodes);
    tmp.rescale_every = (rescale_every > 0 && rescale_every <= 12) ? rescale_every : 8;
    tmp.cherry_tables = cherry_tables != 0;
    return build_schedule(&tmp);
}

static int sr_schedule_dump_impl(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                     const int32_t* leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables, int64_t cap_steps,
                     int32_t* steps, int32_t* node_slot, int64_t* n_steps, int32_t* stack_depth) {
    // host-only: builds the schedule exactly as sr_set_tree does, without touching a GPU
    if (!n_steps) return SR_ERR_ARG;
    sr_ctx tmp;
    int rc = host_schedule(tmp, n_nodes, child_off, child_idx, blen, leaf_row, root, rescale_every, cherry_tables);
    if (rc) return rc;
    *n_steps = (int64_t)tmp.steps.size() / STEP_INTS;
    if (stack_depth) *stack_depth = tmp.need_depth;
    if (steps) {
        if (*n_steps > cap_steps) return SR_ERR_ARG;
        memcpy(steps, tmp.steps.data(), sizeof(int32_t) * tmp.steps.size());
    }
    if (node_slot) memcpy(node_slot, tmp.node_slot.data(), sizeof(int32_t) * n_nodes);
    return SR_OK;
}

static int sr_schedule_records_impl(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                     const int32_t* leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables, int64_t cap_steps,
                     uint32_t* crec, uint32_t* prec, int32_t* clade_gathers, int64_t* n_steps, int64_t* n_clade_gathers) {
    // host-only: the two record streams prune_kernel reads (prune.cuh), for the CPU test that interprets them
    if (!n_steps || !n_clade_gathers || !crec || !prec || !clade_gathers) return SR_ERR_ARG;
    sr_ctx tmp;
    int rc = host_schedule(tmp, n_nodes, child_off, child_idx, blen, leaf_row, root, rescale_every, cherry_tables);
    if (rc) return rc;
    *n_steps = (int64_t)tmp.steps.size() / STEP_INTS;
    *n_clade_gathers = (int64_t)tmp.clade_gathers.size() / 4;
    if (*n_steps > cap_steps || *n_clade_gathers > 3 * cap_steps) return SR_ERR_ARG;
    memcpy(crec, tmp.crec.data(), sizeof(uint32_t) * tmp.crec.size());
    memcpy(prec, tmp.prec.data(), sizeof(uint32_t) * tmp.prec.size());
    memcpy(clade_gathers, tmp.clade_gathers.data(), sizeof(int32_t) * tmp.clade_gathers.size());
    return SR_OK;
}

static int sr_measure_fp64_peak_impl(sr_ctx* c, double ms_target, double* tflops) {
    if (!c || !tflops) return SR_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, c->d_peak.reserve(64));
    const int blocks = c->n_sm * 4, threads = 256, iters = 2048;
    const double flop = 2.0 * 256.0 * (double)(blocks * threads / 32) * iters * 4 * 8;
    cudaEvent_t a, b;
    CUDA_TRY(c, cudaEventCreate(&a)); CUDA_TRY(c, cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) dmma_peak_kernel<<<blocks, threads, 0, c->stream>>>(c->d_peak.as<double>(), iters, 0.5, 0.25);
    double best = 0.0, spent = 0.0;
    while (spent < ms_target) {
        cudaEventRecord(a, c->stream);
        dmma_peak_kernel<<<blocks, threads, 0, c->stream>>>(c->d_peak.as<double>(), iters, 0.5, 0.25);
        cudaEventRecord(b, c->stream);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { cudaEventDestroy(a); cudaEventDestroy(b); return fail(c, SR_ERR_CUDA, "peak probe failed: %s", cudaGetErrorString(e)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        spent += ms;
        best = std::max(best, flop / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops = best;
    return SR_OK;
}

static int sr_debug_pmatrix_impl(sr_ctx* c, int32_t node, int32_t k, double* out) {
    if (!c || !out) return SR_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_pstream(c, c->stream);
    if (rc) return rc;
    if (k < 0 || k >= c->K || node < 0 || node >= c->n_nodes) return fail(c, SR_ERR_ARG, "node or rate class out of range");
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (node == c->root) {
        CUDA_TRY(c, cudaMemcpyAsync(out, c->d_proot.as<double>() + (size_t)k * 400, 3200, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        return SR_OK;
    }
    const int slot = c->node_slot[node];
    if (slot < 0) return fail(c, SR_ERR_ARG, "node %d has no matrix in the stream (root child of a one-tip clade, or part of a collapsed cherry: see sr_debug_cherry_table / option cherry_tables)", node);
    double buf[FRAG_DOUBLES];
    const bool frag = c->slot_kind[slot] == SLOT_FRAG;
    CUDA_TRY(c, cudaMemcpyAsync(buf, c->d_pstream.as<double>() + (size_t)k * c->rate_stride + c->slot_off[slot],
                                frag ? FRAG_BYTES : TABLE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (frag) {
        for (int e = 0; e < FRAG_DOUBLES; ++e) {
            const int half = e & 1, lane = (e >> 1) & 31, f = ((e >> 6) << 1) | half;
            if (f >= 15) continue;
            const int s = f / 3, tt = f % 3, g = lane >> 2, q = lane & 3, r = 2 * tt + (g & 1);
            if (r < 5) out[(5 * (g >> 1) + r) * 20 + 5 * q + s] = buf[e];
        }
    } else {
        for (int s = 0; s < 20; ++s)
            for (int i = 0; i < 20; ++i) out[i * 20 + s] = buf[s * TAB_ROW_DOUBLES + (i / 5) * 6 + i % 5];
    }
    return SR_OK;
}

static int sr_debug_cherry_table_impl(sr_ctx* c, int32_t node, int32_t k, double* out, int64_t cap_doubles, int32_t* n_rows) {
    if (!c || !out) return SR_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_pstream(c, c->stream);
    if (rc) return rc;
    if (k < 0 || k >= c->K || node < 0 || node >= c->n_nodes) return fail(c, SR_ERR_ARG, "node or rate class out of range");
    if (c->node_slot[node] > -2) return fail(c, SR_ERR_ARG, "node %d is not a collapsed clade", node);
    const int cid = -2 - c->node_slot[node];
    const int rows = c->cherry_kind[cid] == 2 ? CHERRY_ROWS : PITCHFORK_ROWS;
    if (n_rows) *n_rows = rows;
    if (cap_doubles < (int64_t)rows * 20) return fail(c, SR_ERR_ARG, "output buffer too small: %d rows of 20 doubles", rows);
    std::vector<double> buf((size_t)rows * CHERRY_ROW_DOUBLES);
    const int64_t stride = c->ctab_rows * CHERRY_ROW_DOUBLES;
    CUDA_TRY(c, cudaMemcpyAsync(buf.data(), c->d_ctab.as<double>() + (size_t)k * stride + (size_t)c->cherry_rowoff[cid] * CHERRY_ROW_DOUBLES,
                                sizeof(double) * buf.size(), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int row = 0; row < rows; ++row)
        for (int i = 0; i < 20; ++i) out[(size_t)row * 20 + i] = buf[(size_t)row * CHERRY_ROW_DOUBLES + (i / 5) * 6 + i % 5];
    return SR_OK;
}

#ifdef SR_TIMELINE
// tools/timeline.py only (not part of the ABI): clock stamps of the last prune launch, [16 warps][SR_TL_STEPS][8]
static int sr_debug_timeline_impl(sr_ctx* c, unsigned* out, int64_t n_words) {
    if (!c || !out || !c->d_timeline.p) return SR_ERR_ARG;
    CUDA_TRY(c, cudaDeviceSynchronize());
    CUDA_TRY(c, cudaMemcpy(out, c->d_timeline.p, sizeof(unsigned) * std::min<int64_t>(n_words, 16 * SR_TL_STEPS * 8), cudaMemcpyDeviceToHost));
    return SR_OK;
}
#endif



// ================================================================================================ one process, N GPUs
}  // namespace

struct sr_multi {
    std::vector<sr_ctx*> ctx;
    std::string err;
};

namespace {

thread_local std::string g_multi_create_error;

int mfail(sr_multi* m, int code, const char* fmt, ...) noexcept {
    char buf[640];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    try {
        if (m) m->err = buf; else g_multi_create_error = buf;
    } catch (...) {
    }
    return code;
}

template <typename F>
int mguarded(sr_multi* m, F&& f) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        return mfail(m, SR_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return mfail(m, SR_ERR_ARG, "internal error: %s", e.what());
    } catch (...) {
        return mfail(m, SR_ERR_ARG, "internal error");
    }
}

// run f(ctx, g) for every context on its own host thread; the first failure (lowest device) is reported
template <typename F>
int multi_each(sr_multi* m, F&& f) {
    const int G = (int)m->ctx.size();
    std::vector<int> rc(G, SR_OK);
    if (G == 1) {
        rc[0] = f(m->ctx[0], 0);
    } else {
        std::vector<std::thread> th;
        th.reserve(G);
        for (int g = 0; g < G; ++g)
            th.emplace_back([&, g] { rc[g] = guarded(m->ctx[g], [&] { return f(m->ctx[g], g); }); });
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < G; ++g)
        if (rc[g]) return mfail(m, rc[g], "device %d: %s", m->ctx[g]->device, m->ctx[g]->err.c_str());
    return SR_OK;
}

int sr_multi_create_impl(sr_multi** out, int n_dev, const int* dev_ids) {
    if (!out) return mfail(nullptr, SR_ERR_ARG, "out is NULL");
    *out = nullptr;
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0) {
        cudaGetLastError();
        return mfail(nullptr, SR_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (n_dev <= 0) { n_dev = visible; dev_ids = nullptr; }
    sr_multi* m = new sr_multi();
    for (int g = 0; g < n_dev; ++g) {
        sr_ctx* c = nullptr;
        const int rc = sr_create(&c, dev_ids ? dev_ids[g] : g);
        if (rc) {
            mfail(nullptr, rc, "%s", sr_last_error(nullptr));
            for (sr_ctx* x : m->ctx) sr_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return SR_OK;
}

int sr_multi_run_streamed_impl(sr_multi* m, int32_t n_taxa, int64_t n_sites, const uint8_t* states, int64_t pitch, int64_t site0,
                               int64_t n, double* lnl, double* joint, sr_compact* compact, double threshold) {
    if (!m) return SR_ERR_ARG;
    if (site0 < 0 || n < 0 || site0 + n > n_sites) return mfail(m, SR_ERR_ARG, "column range outside the alignment");
    const int G = (int)m->ctx.size();
    // contiguous blocks of ceil(n / G) columns; whole 16-column groups so that every block starts aligned
    int64_t per = (n + G - 1) / G;
    per = (per + 15) / 16 * 16;
    return multi_each(m, [&](sr_ctx* c, int g) {
        const int64_t b0 = std::min<int64_t>(n, (int64_t)g * per), b1 = std::min<int64_t>(n, b0 + per);
        if (b1 <= b0) return (int)SR_OK;
        const size_t cstride = (size_t)compact_stride(c->compact_cap);
        return sr_run_streamed_impl(c, n_taxa, n_sites, states, pitch, site0 + b0, b1 - b0, lnl ? lnl + b0 : nullptr,
                                    joint ? joint + b0 * 400 : nullptr,
                                    compact ? reinterpret_cast<sr_compact*>(reinterpret_cast<char*>(compact) + cstride * b0) : nullptr, threshold);
    });
}

}  // namespace

extern "C" {

int sr_compact_cap_for(double threshold) { return compact_cap_for(threshold); }
size_t sr_compact_stride(int cap) { return (cap < 1 || cap > 400) ? 0 : (size_t)compact_stride(cap); }

int sr_multi_create(sr_multi** out, int n_dev, const int* dev_ids) {
    return mguarded(nullptr, [&] { return sr_multi_create_impl(out, n_dev, dev_ids); });
}
void sr_multi_destroy(sr_multi* m) {
    if (!m) return;
    for (sr_ctx* c : m->ctx) sr_destroy(c);
    delete m;
}
const char* sr_multi_last_error(const sr_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
int sr_multi_count(const sr_multi* m) { return m ? (int)m->ctx.size() : 0; }
sr_ctx* sr_multi_ctx(sr_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
int sr_multi_set_model(sr_multi* m, const double* eval, const double* evec, const double* ievc, const double* pi) {
    if (!m) return SR_ERR_ARG;
    return mguarded(m, [&] { return multi_each(m, [&](sr_ctx* c, int) { return sr_set_model_impl(c, eval, evec, ievc, pi); }); });
}
int sr_multi_set_rates(sr_multi* m, int K, const double* rates) {
    if (!m) return SR_ERR_ARG;
    return mguarded(m, [&] { return multi_each(m, [&](sr_ctx* c, int) { return sr_set_rates_impl(c, K, rates); }); });
}
int sr_multi_set_tree(sr_multi* m, int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                      const int32_t* leaf_row, int32_t root) {
    if (!m) return SR_ERR_ARG;
    return mguarded(m, [&] { return multi_each(m, [&](sr_ctx* c, int) { return sr_set_tree_impl(c, n_nodes, child_off, child_idx, blen, leaf_row, root); }); });
}
int sr_multi_set_option(sr_multi* m, const char* name, int64_t value) {
    if (!m) return SR_ERR_ARG;
    return mguarded(m, [&] { return multi_each(m, [&](sr_ctx* c, int) { return sr_set_option_impl(c, name, value); }); });
}
int sr_multi_run_streamed(sr_multi* m, int32_t n_taxa, int64_t n_sites, const uint8_t* states, int64_t pitch, int64_t site0,
                          int64_t n, double* lnl, double* joint, sr_compact* compact, double threshold) {
    if (!m) return SR_ERR_ARG;
    return mguarded(m, [&] { return sr_multi_run_streamed_impl(m, n_taxa, n_sites, states, pitch, site0, n, lnl, joint, compact, threshold); });
}

// ---- the exported symbols: every entry point is fenced, nothing may unwind into a JVM (or any C caller)
void sr_destroy(sr_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& e : c->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (int i = 0; i < 3; ++i) if (c->h_bounce[i]) cudaFreeHost(c->h_bounce[i]);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->ev_stage) cudaEventDestroy(c->ev_stage);
    for (int i = 0; i < 3; ++i) {
        if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
        if (c->ev_sfree[i]) cudaEventDestroy(c->ev_sfree[i]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    delete c;
}

const char* sr_last_error(const sr_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

void* sr_host_alloc(size_t bytes) {
    // portable: page-locked for every device, so one buffer can feed all the contexts of an sr_multi
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void sr_host_free(void* p) { if (p) cudaFreeHost(p); }

#ifdef SR_TIMELINE
int sr_debug_timeline(sr_ctx* c, unsigned* out, int64_t n_words) {
    return guarded(c, [&] { return sr_debug_timeline_impl(c, out, n_words); });
}
#endif

int sr_create(sr_ctx** out, int device) {
    return guarded(nullptr, [&] { return sr_create_impl(out, device); });
}
int sr_set_model(sr_ctx* c, const double* eval, const double* evec, const double* ievc, const double* pi) {
    return guarded(c, [&] { return sr_set_model_impl(c, eval, evec, ievc, pi); });
}
int sr_set_rates(sr_ctx* c, int K, const double* rates) {
    return guarded(c, [&] { return sr_set_rates_impl(c, K, rates); });
}
int sr_set_tree(sr_ctx* c, int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                const int32_t* leaf_row, int32_t root) {
    return guarded(c, [&] { return sr_set_tree_impl(c, n_nodes, child_off, child_idx, blen, leaf_row, root); });
}
int sr_set_alignment(sr_ctx* c, int32_t n_taxa, int64_t n_sites, const uint8_t* states, int64_t pitch) {
    return guarded(c, [&] { return sr_set_alignment_impl(c, n_taxa, n_sites, states, pitch); });
}
int sr_run_device(sr_ctx* c, int64_t site0, int64_t n, double* d_lnl, double* d_joint, sr_compact* d_compact,
                  double threshold, void* stream) {
    return guarded(c, [&] { return sr_run_device_impl(c, site0, n, d_lnl, d_joint, d_compact, threshold, stream); });
}
int sr_run(sr_ctx* c, int64_t site0, int64_t n, double* lnl, double* joint, sr_compact* compact, double threshold) {
    return guarded(c, [&] { return sr_run_impl(c, site0, n, lnl, joint, compact, threshold); });
}
int sr_run_streamed(sr_ctx* c, int32_t n_taxa, int64_t n_sites, const uint8_t* states, int64_t pitch, int64_t site0,
                    int64_t n, double* lnl, double* joint, sr_compact* compact, double threshold) {
    return guarded(c, [&] { return sr_run_streamed_impl(c, n_taxa, n_sites, states, pitch, site0, n, lnl, joint, compact, threshold); });
}
int sr_set_option(sr_ctx* c, const char* name, int64_t value) {
    return guarded(c, [&] { return sr_set_option_impl(c, name, value); });
}
int sr_get_profile(sr_ctx* c, sr_profile* out) {
    return guarded(c, [&] { return sr_get_profile_impl(c, out); });
}
int sr_profile_reset(sr_ctx* c) {
    return guarded(c, [&] { return sr_profile_reset_impl(c); });
}
int sr_schedule_info(sr_ctx* c, int64_t* n_steps, int32_t* stack_depth, double* flops) {
    return guarded(c, [&] { return sr_schedule_info_impl(c, n_steps, stack_depth, flops); });
}
int sr_schedule_dump(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                     const int32_t* leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables, int64_t cap_steps,
                     int32_t* steps, int32_t* node_slot, int64_t* n_steps, int32_t* stack_depth) {
    return guarded(nullptr, [&] { return sr_schedule_dump_impl(n_nodes, child_off, child_idx, blen, leaf_row, root, rescale_every, cherry_tables, cap_steps, steps, node_slot, n_steps, stack_depth); });
}
int sr_schedule_records(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx, const double* blen,
                     const int32_t* leaf_row, int32_t root, int32_t rescale_every, int32_t cherry_tables, int64_t cap_steps,
                     uint32_t* crec, uint32_t* prec, int32_t* clade_gathers, int64_t* n_steps, int64_t* n_clade_gathers) {
    return guarded(nullptr, [&] { return sr_schedule_records_impl(n_nodes, child_off, child_idx, blen, leaf_row, root, rescale_every, cherry_tables, cap_steps, crec, prec, clade_gathers, n_steps, n_clade_gathers); });
}
int sr_measure_fp64_peak(sr_ctx* c, double ms_target, double* tflops) {
    return guarded(c, [&] { return sr_measure_fp64_peak_impl(c, ms_target, tflops); });
}
int sr_debug_pmatrix(sr_ctx* c, int32_t node, int32_t k, double* out) {
    return guarded(c, [&] { return sr_debug_pmatrix_impl(c, node, k, out); });
}
int sr_debug_cherry_table(sr_ctx* c, int32_t node, int32_t k, double* out, int64_t cap_doubles, int32_t* n_rows) {
    return guarded(c, [&] { return sr_debug_cherry_table_impl(c, node, k, out, cap_doubles, n_rows); });
}

}  // extern "C"
