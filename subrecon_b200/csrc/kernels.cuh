// kernels.cuh — device code of the SubRecon hot path for sm_100a (B200).
//
//   K1  pbuild_kernel   P(rt) = |Evec · diag(exp(λ r t)) · Ievc| per (pruning step, rate class), written
//                       straight into the layout the pruning kernel consumes (MMA B-fragments or a
//                       transposed gather table). Reference: PAL MatrixExponential::setDistance,
//                       called at recon/JointBranchReconstruction.java:99,216.
//   K2  prune_kernel    post-order Felsenstein pruning of both root clades for a tile of columns and one
//                       rate class per work item (JointBranchReconstruction.java:191-247). The 20x20 · 20xS
//                       contraction runs on the FP64 tensor pipe (DMMA m8n8k4); the per-step matrices are
//                       streamed through a shared-memory ring by 1-D TMA bulk copies (cp.async.bulk +
//                       mbarrier); partials never leave the SM (registers + a shared-memory stack).
//   K3  root_kernel     400 joint probabilities + marginal lnL per column over K rate classes
//                       (JointBranchReconstruction.java:99-139, utils/Utils.java:47-80).
//
// State layout inside K2. A warp owns M "column tiles" of 8 alignment columns. For tile j, lane = 4g+q
// (g = lane>>2 the column inside the tile, q = lane&3) holds the 5 conditional likelihoods of states
// 5q..5q+4 in A[j][0..4]. With that choice the same registers serve as the DMMA A-operand (row g,
// k-slot q of k-step s  <->  state 5q+s) and as its D-result (row g, n-slot 8t+2q+e  <->  state 5q+2t+e,
// the slot t=2,e=1 is padding), so the tree walk needs no shuffles between levels.
#pragma once
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

namespace srk {

constexpr int NS = 20;
constexpr int SLOT_DOUBLES = 480;                 // one pruning step's matrix in either layout
constexpr int SLOT_BYTES = SLOT_DOUBLES * 8;      // 3840
constexpr int LEAF_TABLE_BYTES = 21 * NS * 8;     // 3360: P^T rows for states 0..19 + the missing-data row
constexpr int STAGE_BYTES = 4096;                 // slot + up to 256 state bytes
constexpr int N_CONSUMER_WARPS = 8;
constexpr int PRUNE_THREADS = (N_CONSUMER_WARPS + 1) * 32;

// op word of the pruning schedule
enum : int32_t {
    OP_LEAF_SET = 0,        // A  = T[state]
    OP_LEAF_MUL = 1,        // A *= T[state]
    OP_MV_SET = 2,          // A  = P·A
    OP_MV_MULPOP = 3,       // A  = pop() ∘ (P·A)
    OP_CODE_MASK = 3,
    OPF_RESCALE = 1 << 2,   // after the op: scale A by a power of two, count the exponent
    OPF_PUSH = 1 << 3,      // after the op (and rescale): push A
    OPF_STORE_A = 1 << 4,   // after the op: A is the partial of root child 0
    OPF_STORE_B = 1 << 5,   // ... of root child 1 (also stores the exponent count)
    OPF_IDENTITY = 1 << 6,  // K1 only: tip directly below the root, table = identity
    OP_ROW_SHIFT = 8        // leaf ops: alignment row in bits 8..31
};

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Consumer release of a ring stage that was read with ordinary ld.shared: the reads (generic proxy) must be
// ordered before the producer's next TMA write (async proxy) to the same bytes. Without this fence the arrive can
// overtake loads still queued in the LSU and a lagging warp reads a half-overwritten matrix (seen on B200 as
// run-to-run 1e-12..1e-2 relative noise in the warps sharing an SMSP with the producer).
__device__ __forceinline__ void stage_release(uint64_t* bar, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// FP64 tensor-core MMA, D(8x8) += A(8x4) · B(4x8)   (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------- K1
struct PBuildParams {
    const double* eval;      // [20]
    const double* evec;      // [400]  Evec[i][k]
    const double* ievc;      // [400]  Ievc[k][j]
    const double* rates;     // [K]
    const int32_t* ops;      // [n_ops]
    const double* op_blen;   // [n_ops] branch length of the edge the step applies
    double* pstream;         // [K][n_ops][480]
    double* proot;           // [K][400] row-major, merged root branch
    double root_blen;        // tA + tB (summed before the rate multiply, JointBranchReconstruction.java:99)
    int n_ops;
    int K;
};

// grid (n_ops + 1, K); block 128. Block x == n_ops builds the root-branch matrix.
__global__ void __launch_bounds__(128) pbuild_kernel(PBuildParams p) {
    __shared__ double iexp[NS * NS];
    __shared__ double P[NS * NS];
    const int op = blockIdx.x, k = blockIdx.y, tid = threadIdx.x;
    const bool is_root = (op == p.n_ops);
    const int32_t w = is_root ? 0 : p.ops[op];
    const double rate = p.rates[k];
    // same expression order as the reference: child.getBranchLength() * rate  /  (tA + tB) * rate
    double t = (is_root ? p.root_blen : p.op_blen[op]) * rate;
    if (t < 1e-9) t = 1e-9;                                        // MatrixExponential::setDistance @0-11
    for (int e = tid; e < NS * NS; e += 128) {
        const int kk = e / NS;
        iexp[e] = __dmul_rn(p.ievc[e], exp(__dmul_rn(t, p.eval[kk])));
    }
    __syncthreads();
    for (int e = tid; e < NS * NS; e += 128) {
        const int i = e / NS, j = e % NS;
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < NS; ++kk) s = __dadd_rn(s, __dmul_rn(p.evec[i * NS + kk], iexp[kk * NS + j]));
        P[e] = fabs(s);                                            // @85-176
    }
    __syncthreads();
    if (is_root) {
        for (int e = tid; e < NS * NS; e += 128) p.proot[(size_t)k * NS * NS + e] = P[e];
        return;
    }
    double* out = p.pstream + ((size_t)k * p.n_ops + op) * SLOT_DOUBLES;
    const int code = w & OP_CODE_MASK;
    if (code >= OP_MV_SET) {
        // B-fragments: slot (s*3+t)*32 + lane = P[5(g/2) + 2t + (g&1)][5q + s], g = lane>>2, q = lane&3
        for (int e = tid; e < SLOT_DOUBLES; e += 128) {
            const int lane = e & 31, st = e >> 5, s = st / 3, tt = st % 3;
            const int g = lane >> 2, q = lane & 3;
            const int r = 2 * tt + (g & 1);
            out[e] = (r < 5) ? P[(5 * (g >> 1) + r) * NS + 5 * q + s] : 0.0;
        }
    } else {
        // gather table T[s][i] = P[i][s]; row 20 = Σ_j P[i][j] (a missing tip is the all-ones vector,
        // JointBranchReconstruction.java:201-205, so its contribution is the row sum, left to right)
        const bool ident = (w & OPF_IDENTITY) != 0;
        for (int e = tid; e < SLOT_DOUBLES; e += 128) {
            double v = 0.0;
            if (e < 21 * NS) {
                const int s = e / NS, i = e % NS;
                if (s < NS) v = ident ? (i == s ? 1.0 : 0.0) : P[i * NS + s];
                else if (ident) v = 1.0;
                else {
                    double acc = 0.0;
                    for (int j = 0; j < NS; ++j) acc = __dadd_rn(acc, P[i * NS + j]);
                    v = acc;
                }
            }
            out[e] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------- K2
struct PruneParams {
    const int32_t* ops;        // [n_ops]
    const double* pstream;     // [K][n_ops][480]
    const uint8_t* states;     // device alignment block, states[row*pitch + col]
    int64_t pitch;
    int64_t col0;              // first column of this launch inside the block (multiple of 16)
    double* rootpart;          // [2][K][ncp][20]
    int32_t* rootexp;          // [K][ncp]
    double* spill;             // global overflow of the partial stack
    int64_t ncp;               // padded column count of this launch (n_tiles * S)
    int n_ops;
    int n_tiles;
    int K;
    int smem_levels;           // stack levels resident in shared memory
    int spill_levels;
};

template <int M>
struct PruneCfg {
    static constexpr int W = N_CONSUMER_WARPS;
    static constexpr int S = W * M * 8;                 // columns per work item
    static constexpr int LEVEL_DOUBLES = W * M * 5 * 32;
    static constexpr int LEVEL_BYTES = LEVEL_DOUBLES * 8;
};

__host__ __device__ constexpr int prune_bar_bytes(int nstage) { return ((2 * nstage * 8 + 127) / 128) * 128; }

template <int M, int NSTAGE>
__global__ void __launch_bounds__(PRUNE_THREADS, 1) prune_kernel(PruneParams p) {
    using Cfg = PruneCfg<M>;
    constexpr int W = Cfg::W;
    constexpr int S = Cfg::S;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stages = smem;                                             // NSTAGE * STAGE_BYTES
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty = full + NSTAGE;
    double* stack = reinterpret_cast<double*>(smem + NSTAGE * STAGE_BYTES + prune_bar_bytes(NSTAGE));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], W); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_items = p.n_tiles * p.K;
    uint32_t it = 0;                                  // ring position, runs on across work items

    if (warp == W) {
        // ===== producer: one lane streams every step's matrix (and tip states) into the ring =====
        if (lane == 0) {
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int k = item / p.n_tiles, tile = item - k * p.n_tiles;
                const double* ps = p.pstream + (size_t)k * p.n_ops * SLOT_DOUBLES;
                const uint8_t* st = p.states + p.col0 + (int64_t)tile * S;
                for (int op = 0; op < p.n_ops; ++op, ++it) {
                    const int stage = it % NSTAGE;
                    const uint32_t ph = (it / NSTAGE) & 1;
                    const int32_t w = p.ops[op];
                    mbar_wait(&empty[stage], ph ^ 1);
                    unsigned char* dst = stages + stage * STAGE_BYTES;
                    if ((w & OP_CODE_MASK) >= OP_MV_SET) {
                        mbar_expect_tx(&full[stage], SLOT_BYTES);
                        bulk_g2s(dst, ps + (size_t)op * SLOT_DOUBLES, SLOT_BYTES, &full[stage]);
                    } else {
                        mbar_expect_tx(&full[stage], LEAF_TABLE_BYTES + S);
                        bulk_g2s(dst, ps + (size_t)op * SLOT_DOUBLES, LEAF_TABLE_BYTES, &full[stage]);
                        bulk_g2s(dst + SLOT_BYTES, st + (int64_t)(uint32_t(w) >> OP_ROW_SHIFT) * p.pitch, S, &full[stage]);
                    }
                }
            }
        }
        return;
    }

    // ===== consumers: 8 warps, each walks the whole schedule for its own 8*M columns =====
    const int g = lane >> 2, q = lane & 3;
    double* spill_base = p.spill + ((size_t)blockIdx.x * W + warp) * (size_t)p.spill_levels * (M * 5 * 32);

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int k = item / p.n_tiles, tile = item - k * p.n_tiles;
        double A[M][5];
        int cnt[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
            cnt[j] = 0;
#pragma unroll
            for (int r = 0; r < 5; ++r) A[j][r] = 0.0;
        }
        int sp = 0;
        int32_t w_next = p.ops[0];
        for (int op = 0; op < p.n_ops; ++op, ++it) {
            const int stage = it % NSTAGE;
            const uint32_t ph = (it / NSTAGE) & 1;
            const int32_t w = w_next;
            if (op + 1 < p.n_ops) w_next = p.ops[op + 1];
            mbar_wait(&full[stage], ph);
            const double* sd = reinterpret_cast<const double*>(stages + stage * STAGE_BYTES);
            const int code = w & OP_CODE_MASK;
            if (code >= OP_MV_SET) {
                double Bf[15];
#pragma unroll
                for (int i = 0; i < 15; ++i) Bf[i] = sd[i * 32 + lane];
                double pv[M][5];
                if (code == OP_MV_MULPOP) {
                    --sp;
                    const double* src = (sp < p.smem_levels)
                        ? stack + ((size_t)sp * W + warp) * (M * 5 * 32)
                        : spill_base + (size_t)(sp - p.smem_levels) * (M * 5 * 32);
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) pv[j][r] = src[(j * 5 + r) * 32 + lane];
                }
                double acc[M][3][2];
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int t = 0; t < 3; ++t) { acc[j][t][0] = 0.0; acc[j][t][1] = 0.0; }
#pragma unroll
                for (int s = 0; s < 5; ++s)
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int t = 0; t < 3; ++t) dmma(acc[j][t][0], acc[j][t][1], A[j][s], Bf[s * 3 + t]);
                stage_release(&empty[stage], lane);      // after the MMAs were issued: their operands have left smem
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        const double o = acc[j][r >> 1][r & 1];
                        A[j][r] = (code == OP_MV_MULPOP) ? o * pv[j][r] : o;
                    }
            } else {
                const unsigned char* st = stages + stage * STAGE_BYTES + SLOT_BYTES + warp * (M * 8) + g;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    int s = st[j * 8];
                    s = s > NS ? NS : s;
                    const double* row = sd + s * NS + q * 5;
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        const double v = row[r];
                        A[j][r] = (code == OP_LEAF_MUL) ? A[j][r] * v : v;
                    }
                }
                stage_release(&empty[stage], lane);
            }
            if (w & OPF_RESCALE) {
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    double mx = fmax(fmax(fmax(A[j][0], A[j][1]), fmax(A[j][2], A[j][3])), A[j][4]);
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    const int e = (__double2hiint(mx) >> 20) & 0x7ff;
                    const double sc = __hiloint2double((2046 - e) << 20, 0);      // 2^(1023-e), exact
#pragma unroll
                    for (int r = 0; r < 5; ++r) A[j][r] *= sc;
                    cnt[j] += e - 1023;
                }
            }
            if (w & OPF_PUSH) {
                double* dst = (sp < p.smem_levels)
                    ? stack + ((size_t)sp * W + warp) * (M * 5 * 32)
                    : spill_base + (size_t)(sp - p.smem_levels) * (M * 5 * 32);
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int r = 0; r < 5; ++r) dst[(j * 5 + r) * 32 + lane] = A[j][r];
                ++sp;
            }
            if (w & (OPF_STORE_A | OPF_STORE_B)) {
                const int which = (w & OPF_STORE_B) ? 1 : 0;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    const int64_t col = (int64_t)tile * S + warp * (M * 8) + j * 8 + g;
                    double* dst = p.rootpart + (((size_t)which * p.K + k) * p.ncp + col) * NS + q * 5;
#pragma unroll
                    for (int r = 0; r < 5; ++r) dst[r] = A[j][r];
                    if (which == 1 && q == 0) p.rootexp[(size_t)k * p.ncp + col] = cnt[j];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- K3
struct RootParams {
    const double* rootpart;    // [2][K][ncp][20]
    const int32_t* rootexp;    // [K][ncp]
    const double* proot;       // [K][400]
    const double* pi;          // [20]
    double* lnl;               // [n] or null
    double* joint;             // [n][400] or null
    void* compact;             // sr_compact[n] or null
    double threshold;
    int64_t ncp;
    int64_t n;                 // output columns of this launch
    int64_t col_skip;          // launch column of output 0 (launches start on a 16-column boundary)
    int64_t out0;              // offset of this launch's first output in the output arrays
    int K;
};

struct CompactRec {            // must match sr_compact in include/subrecon_b200.h
    double lnl, max_ii, max_p;
    int32_t n_kept, interesting;
    int16_t pair[8];
    double prob[8];
};

// one warp per column; lane a < 20 owns row a of the 20x20 joint matrix
constexpr int ROOT_WARPS = 8;
__global__ void __launch_bounds__(ROOT_WARPS * 32) root_kernel(RootParams p) {
    __shared__ double sB[ROOT_WARPS][NS];
    __shared__ double sOut[ROOT_WARPS][NS * NS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double LN2 = 0.693147180559945309417232121458;
    for (int64_t oc = (int64_t)blockIdx.x * ROOT_WARPS + warp; oc < p.n; oc += (int64_t)gridDim.x * ROOT_WARPS) {
        const int64_t col = oc + p.col_skip;
        int emax = INT_MIN;
        for (int k = 0; k < p.K; ++k) emax = max(emax, p.rootexp[(size_t)k * p.ncp + col]);
        double row[NS];
#pragma unroll
        for (int b = 0; b < NS; ++b) row[b] = 0.0;
        const double pia = lane < NS ? p.pi[lane] : 0.0;
        for (int k = 0; k < p.K; ++k) {
            const int e = p.rootexp[(size_t)k * p.ncp + col];
            const double wk = ldexp(1.0, e - emax);                                     // exact, <= 1
            const double* pa = p.rootpart + (((size_t)0 * p.K + k) * p.ncp + col) * NS;
            const double* pb = p.rootpart + (((size_t)1 * p.K + k) * p.ncp + col) * NS;
            __syncwarp();
            if (lane < NS) sB[warp][lane] = pb[lane];
            __syncwarp();
            if (lane < NS) {
                const double fa = wk * (pia * pa[lane]);
                const double* pr = p.proot + (size_t)k * NS * NS + lane * NS;
#pragma unroll
                for (int b = 0; b < NS; ++b) row[b] = fma(fa * pr[b], sB[warp][b], row[b]);
            }
        }
        double rs = 0.0;
#pragma unroll
        for (int b = 0; b < NS; ++b) rs += row[b];
        double tot = rs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        const double inv = 1.0 / tot;
        const double lnl = log(tot) + (double)emax * LN2 - log((double)p.K);
        if (p.lnl && lane == 0) p.lnl[p.out0 + oc] = lnl;
        if (p.joint || p.compact) {
            __syncwarp();
            if (lane < NS) {
#pragma unroll
                for (int b = 0; b < NS; ++b) sOut[warp][lane * NS + b] = row[b] * inv;
            }
            __syncwarp();
            if (p.joint) {
                double* dst = p.joint + (size_t)(p.out0 + oc) * (NS * NS);
                for (int e = lane; e < NS * NS; e += 32) dst[e] = sOut[warp][e];
            }
            if (p.compact) {
                // SiteResult's scan (recon/SiteResult.java:69-83): maxII, max, entries >= threshold in a-major order
                double mxp = 0.0, mxii = 0.0;
                int nk = 0;
                for (int e = lane; e < NS * NS; e += 32) {
                    const double v = sOut[warp][e];
                    mxp = fmax(mxp, v);
                    if (e / NS == e % NS) mxii = fmax(mxii, v);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mxp = fmax(mxp, __shfl_xor_sync(0xffffffffu, mxp, o));
                    mxii = fmax(mxii, __shfl_xor_sync(0xffffffffu, mxii, o));
                }
                CompactRec* rec = reinterpret_cast<CompactRec*>(p.compact) + (p.out0 + oc);
                for (int base = 0; base < NS * NS; base += 32) {
                    const int e = base + lane;
                    const bool keep = e < NS * NS && sOut[warp][e] >= p.threshold;
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    const int pos = nk + __popc(m & ((1u << lane) - 1));
                    if (keep && pos < 8) { rec->pair[pos] = (int16_t)e; rec->prob[pos] = sOut[warp][e]; }
                    nk += __popc(m);
                }
                if (lane == 0) {
                    rec->lnl = lnl; rec->max_ii = mxii; rec->max_p = mxp; rec->n_kept = nk;
                    rec->interesting = (mxii <= 1.0 - p.threshold && mxp >= p.threshold) ? 1 : 0;
                }
                if (lane >= nk && lane < 8) { rec->pair[lane] = -1; rec->prob[lane] = 0.0; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- FP64 probe
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = i; c1[i] = -i; }
    for (int itx = 0; itx < iters; ++itx) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

}  // namespace srk
