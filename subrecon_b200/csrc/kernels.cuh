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
constexpr int FRAG_DOUBLES = 512;                 // a 20x20 matrix as 15 (+1 pad) DMMA B-fragments x 32 lanes, stored as
                                                  // 8 pairs x 32 lanes x 2 doubles so a lane fetches two fragments per LDS.128
constexpr int FRAG_BYTES = FRAG_DOUBLES * 8;      // 4096
constexpr int TABLE_DOUBLES = 21 * NS;            // P^T rows for states 0..19 + the missing-data row
constexpr int TABLE_BYTES = TABLE_DOUBLES * 8;    // 3360
constexpr int CHERRY_ROWS = 21 * 21;          // state pairs (s_a, s_b), 20 = missing
constexpr int PITCHFORK_ROWS = 21 * 21 * 21;  // state triples (s_a, s_b, s_t)
constexpr int CHERRY_ROW_DOUBLES = 24;        // 4 groups of 5 values + 1 pad: a lane fetches its group with three 16-byte loads

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
// non-blocking probe (try_wait may suspend the thread for a hardware time slice when the phase is still open)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------- K1
// One "slot" = one transition matrix in the layout its consumer wants.
enum : int32_t { SLOT_FRAG = 0, SLOT_TABLE = 1, SLOT_IDENTITY = 2 };

struct PBuildParams {
    const double* eval;        // [20]
    const double* evec;        // [400]  Evec[i][k]
    const double* ievc;        // [400]  Ievc[k][j]
    const double* rates;       // [K]
    const int32_t* slot_kind;  // [n_slots]
    const int32_t* slot_off;   // [n_slots] offset (doubles) inside one rate class's stream
    const double* slot_blen;   // [n_slots] branch length of the edge
    double* pstream;           // [K][rate_stride]
    double* proot;             // [K][400] row-major, merged root branch
    int64_t rate_stride;       // doubles per rate class
    double root_blen;          // tA + tB (summed before the rate multiply, JointBranchReconstruction.java:99)
    int n_slots;
    int K;
};

// grid (n_slots + 1, K); block 128. Block x == n_slots builds the root-branch matrix.
__global__ void __launch_bounds__(128) pbuild_kernel(PBuildParams p) {
    __shared__ double iexp[NS * NS];
    __shared__ double P[NS * NS];
    const int slot = blockIdx.x, k = blockIdx.y, tid = threadIdx.x;
    const bool is_root = (slot == p.n_slots);
    const int kind = is_root ? -1 : p.slot_kind[slot];
    const double rate = p.rates[k];
    // same expression order as the reference: child.getBranchLength() * rate  /  (tA + tB) * rate
    double t = (is_root ? p.root_blen : p.slot_blen[slot]) * rate;
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
    double* out = p.pstream + (size_t)k * p.rate_stride + p.slot_off[slot];
    if (kind == SLOT_FRAG) {
        // B-fragment f = s*3+t of lane (g = lane>>2, q = lane&3) is P[5(g/2) + 2t + (g&1)][5q + s]; it is stored at
        // ((f>>1)*32 + lane)*2 + (f&1)
        for (int e = tid; e < FRAG_DOUBLES; e += 128) {
            const int half = e & 1, lane = (e >> 1) & 31, f = ((e >> 6) << 1) | half;
            double v = 0.0;
            if (f < 15) {
                const int s = f / 3, tt = f % 3, g = lane >> 2, q = lane & 3, r = 2 * tt + (g & 1);
                if (r < 5) v = P[(5 * (g >> 1) + r) * NS + 5 * q + s];
            }
            out[e] = v;
        }
    } else {
        // gather table T[s][i] = P[i][s]; row 20 = Σ_j P[i][j] (a missing tip is the all-ones vector,
        // JointBranchReconstruction.java:201-205, so its contribution is the row sum, left to right)
        const bool ident = (kind == SLOT_IDENTITY);
        for (int e = tid; e < TABLE_DOUBLES; e += 128) {
            const int s = e / NS, i = e % NS;
            double v;
            if (s < NS) v = ident ? (i == s ? 1.0 : 0.0) : P[i * NS + s];
            else if (ident) v = 1.0;
            else {
                double acc = 0.0;
                for (int j = 0; j < NS; ++j) acc = __dadd_rn(acc, P[i * NS + j]);
                v = acc;
            }
            out[e] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------- K1b
// Cherry tables (see build_schedule): for a collapsed cherry v = (a, b) and every state pair (s_a, s_b) in 0..20 (20 =
// missing data), the vector v hands to its parent, in the reference's own order of operations
// (JointBranchReconstruction.java:208-230): partial_v[k] = P_a[k][s_a] * P_b[k][s_b] (a missing tip contributes the row
// sum of its matrix), then sum_i = sum over k ascending of P_v[i][k] * partial_v[k], separate multiply and add.
// Row layout: 4 groups of {5 values, 1 pad} so that the lane owning states 5q..5q+4 fetches its group with three
// 16-byte loads. grid (n_cherry, K), block 256.
struct CherryParams {
    const double* eval;        // [20]
    const double* evec;        // [400]
    const double* ievc;        // [400]
    const double* rates;       // [K]
    const double* blen;        // [n][5] = {tip a, tip b, top node, (pitchfork:) cherry edge, third tip}
    const int32_t* kind;       // [n] 2 = cherry, 3 = pitchfork
    const int64_t* rowoff;     // [n] first row inside one rate class's block
    double* ctab;              // [K][rows][CHERRY_ROW_DOUBLES]
    int64_t rate_stride;       // doubles per rate class
};

// Dynamic shared memory (pitchforks only): U[441][20], the absorbed cherry's vector per state pair. grid (n, K, slices).
__global__ void __launch_bounds__(256) cherry_kernel(CherryParams p) {
    __shared__ double iexp[NS * NS];
    __shared__ double Pm[5][NS * NS];
    __shared__ double Tt[3][21 * NS];                  // Tt[m][s][k] = P_m[k][s] for tips a, b, t; s = 20: row sum of P_m
    extern __shared__ __align__(16) double U[];
    const int cid = blockIdx.x, kr = blockIdx.y, tid = threadIdx.x;
    const int slice = blockIdx.z, stride = 256 * gridDim.z;     // the table's rows are split over gridDim.z CTAs
    const int kind = p.kind[cid];
    const double rate = p.rates[kr];
    for (int m = 0; m < (kind == 3 ? 5 : 3); ++m) {
        double t = p.blen[cid * 5 + m] * rate;
        if (t < 1e-9) t = 1e-9;
        __syncthreads();
        for (int e = tid; e < NS * NS; e += 256) iexp[e] = __dmul_rn(p.ievc[e], exp(__dmul_rn(t, p.eval[e / NS])));
        __syncthreads();
        for (int e = tid; e < NS * NS; e += 256) {
            const int i = e / NS, j = e % NS;
            double s = 0.0;
#pragma unroll
            for (int kk = 0; kk < NS; ++kk) s = __dadd_rn(s, __dmul_rn(p.evec[i * NS + kk], iexp[kk * NS + j]));
            Pm[m][e] = fabs(s);
        }
    }
    __syncthreads();
    // tip tables of a, b (matrices 0, 1) and, for a pitchfork, t (matrix 4)
    for (int e = tid; e < 3 * 21 * NS; e += 256) {
        const int m = e / (21 * NS), sk = e % (21 * NS), st = sk / NS, k = sk % NS;
        if (m == 2 && kind != 3) continue;
        const double* P = Pm[m == 2 ? 4 : m];
        double v;
        if (st < NS) v = P[k * NS + st];
        else {
            double acc = 0.0;
            for (int j = 0; j < NS; ++j) acc = __dadd_rn(acc, P[k * NS + j]);
            v = acc;
        }
        Tt[m][sk] = v;
    }
    __syncthreads();
    double* out = p.ctab + (size_t)kr * p.rate_stride + (size_t)p.rowoff[cid] * CHERRY_ROW_DOUBLES;
    // One thread per table row: the row's operand w[k] (tip product, or cherry vector x third tip) is formed once and kept
    // in registers; the top node's matrix is then read with warp-uniform (broadcast) addresses. Same operation order as
    // the reference for every entry: sum over k ascending of P[i][k] * w[k], separate multiply and add.
    auto emit_row = [&](int row, const double (&w)[NS]) {
        double* o = out + (size_t)row * CHERRY_ROW_DOUBLES;
#pragma unroll 1
        for (int grp = 0; grp < 4; ++grp) {
            double v[6];
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const double* pv = Pm[2] + (grp * 5 + r) * NS;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < NS; ++k) s = __dadd_rn(s, __dmul_rn(pv[k], w[k]));
                v[r] = s;
            }
            v[5] = 0.0;
#pragma unroll
            for (int r = 0; r < 6; r += 2) *reinterpret_cast<double2*>(o + grp * 6 + r) = make_double2(v[r], v[r + 1]);
        }
    };
    if (kind == 2) {
        for (int row = tid + slice * 256; row < CHERRY_ROWS; row += stride) {
            const double* ta = Tt[0] + (row / 21) * NS;
            const double* tb = Tt[1] + (row % 21) * NS;
            double w[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) w[k] = __dmul_rn(ta[k], tb[k]);
            emit_row(row, w);
        }
        return;
    }
    // pitchfork: first the absorbed cherry's vector up its own edge (matrix 3), per state pair ...
    for (int e = tid; e < CHERRY_ROWS * NS; e += 256) {
        const int row = e / NS, k = e % NS, sa = row / 21, sb = row % 21;
        const double* ta = Tt[0] + sa * NS;
        const double* tb = Tt[1] + sb * NS;
        const double* pc = Pm[3] + k * NS;
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < NS; ++m) s = __dadd_rn(s, __dmul_rn(pc[m], __dmul_rn(ta[m], tb[m])));
        U[e] = s;
    }
    __syncthreads();
    // ... then the top node: (cherry vector x third tip) carried up the node's own edge (matrix 2)
    for (int row = tid + slice * 256; row < PITCHFORK_ROWS; row += stride) {
        const double* u = U + (row / 21) * NS;
        const double* tt = Tt[2] + (row % 21) * NS;
        double w[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) w[k] = __dmul_rn(u[k], tt[k]);
        emit_row(row, w);
    }
}

// ------------------------------------------------------------------------------------------- K0
// In-place clamp of the uploaded state codes to 0..20 (20 = missing data: any code outside 0..19,
// JointBranchReconstruction.java:199-205), so the pruning kernel can index its 21-row gather tables directly.
__global__ void __launch_bounds__(256) sanitize_kernel(uint4* __restrict__ v, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 x = v[i];
        x.x = __vminu4(x.x, 0x14141414u); x.y = __vminu4(x.y, 0x14141414u);
        x.z = __vminu4(x.z, 0x14141414u); x.w = __vminu4(x.w, 0x14141414u);
        v[i] = x;
    }
}

// Same pass for RESIDUE CHARACTERS ("input_chars"): the bytes are the alignment's letters as PAL holds them
// (`SimpleAlignment.getData` = `String.charAt`), translated the way `AminoAcids::getStateImpl` does:
// letters of either case A0 R1 N2 D3 C4 Q5 E6 G7 H8 I9 L10 K11 M12 F13 P14 S15 T16 W17 Y18 V19; gaps, '?', '*',
// B J O U X Z and every other byte are outside 0..19 for the reference and therefore missing data (20) here. This is
// the reference's per-cell `getStateBySequenceName` (molevo/AdvancedAlignmentAminoAcid.java:34-40) as one streaming pass.
__global__ void __launch_bounds__(256) encode_chars_kernel(uint4* __restrict__ v, size_t n16) {
    __shared__ uint8_t lut[256];
    {
        //                      A  B   C  D  E  F   G  H  I  J   K   L   M   N  O   P   Q  R  S   T   U   V   W   X   Y   Z
        const uint8_t pos[26] = {0, 20, 4, 3, 6, 13, 7, 8, 9, 20, 11, 10, 12, 2, 20, 14, 5, 1, 15, 16, 20, 19, 17, 20, 18, 20};
        const int c = threadIdx.x, lower = c | 0x20;
        const bool letter = (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
        lut[c] = letter ? pos[lower - 'a'] : (uint8_t)20;
    }
    __syncthreads();
    auto tr = [&](unsigned x) {
        return (unsigned)lut[x & 0xff] | ((unsigned)lut[(x >> 8) & 0xff] << 8) | ((unsigned)lut[(x >> 16) & 0xff] << 16) |
               ((unsigned)lut[x >> 24] << 24);
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 x = v[i];
        x.x = tr(x.x); x.y = tr(x.y); x.z = tr(x.z); x.w = tr(x.w);
        v[i] = x;
    }
}

// ------------------------------------------------------------------------------------------- column patterns
// Identical alignment columns have identical results, so a batch can be reduced to its distinct column patterns
// (the reference recomputes every column, SubRecon.java:112-116). A pattern is identified by a 128-bit
// hash of its state codes; equal hashes are treated as equal columns.
__device__ __forceinline__ uint64_t mix64(uint64_t x) {          // murmur3 finaliser
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256) hash_columns_kernel(const uint8_t* __restrict__ states, int64_t pitch, int n_taxa,
                                                           int64_t n, uint64_t* __restrict__ h1, uint64_t* __restrict__ h2,
                                                           uint32_t* __restrict__ idx) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    uint64_t a = 0x9e3779b97f4a7c15ULL, b = 0xc2b2ae3d27d4eb4fULL;
    for (int t = 0; t < n_taxa; ++t) {                           // adjacent threads read adjacent bytes of one row
        const uint64_t v = states[(int64_t)t * pitch + c];
        a = (a ^ v) * 0x100000001b3ULL;
        b = mix64(b + v + 0x632be59bd9b4e019ULL * (uint64_t)(t + 1));
    }
    h1[c] = mix64(a);
    h2[c] = b;
    idx[c] = (uint32_t)c;
}
// after sorting (h1, idx): flag the first column of every run of equal (h1, h2)
__global__ void __launch_bounds__(256) pattern_heads_kernel(const uint64_t* __restrict__ h1s, const uint32_t* __restrict__ idxs,
                                                            const uint64_t* __restrict__ h2, int64_t n, int32_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || h1s[i] != h1s[i - 1] || h2[idxs[i]] != h2[idxs[i - 1]]) ? 1 : 0;
}
// uid = inclusive scan of head: pattern id (1-based) of the i-th sorted column
__global__ void __launch_bounds__(256) pattern_map_kernel(const uint32_t* __restrict__ idxs, const int32_t* __restrict__ head,
                                                          const int32_t* __restrict__ uid, int64_t n, uint32_t* __restrict__ rep,
                                                          uint32_t* __restrict__ map) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t u = (uint32_t)uid[i] - 1;
    map[idxs[i]] = u;
    if (head[i]) rep[u] = idxs[i];
}
__global__ void __launch_bounds__(256) gather_patterns_kernel(const uint8_t* __restrict__ states, int64_t pitch, int n_taxa,
                                                              const uint32_t* __restrict__ rep, int64_t n_uniq,
                                                              uint8_t* __restrict__ out, int64_t out_pitch) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (u < out_pitch && t < n_taxa) out[(int64_t)t * out_pitch + u] = (u < n_uniq) ? states[(int64_t)t * pitch + rep[u]] : (uint8_t)NS;
}
// results of the distinct patterns back to every original column: one warp per column
__global__ void __launch_bounds__(256) scatter_patterns_kernel(const uint32_t* __restrict__ map, int64_t n, int64_t out0,
                                                               const double* __restrict__ u_lnl, const double* __restrict__ u_joint,
                                                               const uint4* __restrict__ u_compact, double* __restrict__ lnl,
                                                               double* __restrict__ joint, uint4* __restrict__ compact) {
    const int lane = threadIdx.x & 31;
    for (int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const uint32_t u = map[c];
        if (lnl && lane == 0) lnl[out0 + c] = u_lnl[u];
        if (joint) {
            const double2* src = reinterpret_cast<const double2*>(u_joint + (size_t)u * 400);
            double2* dst = reinterpret_cast<double2*>(joint + (size_t)(out0 + c) * 400);
            for (int e = lane; e < 200; e += 32) dst[e] = src[e];
        }
        if (compact && lane < 7) compact[(size_t)(out0 + c) * 7 + lane] = u_compact[(size_t)u * 7 + lane];   // 112 B = 7 x 16 B
    }
}

// ------------------------------------------------------------------------------------------- K2
// A pruning STEP (one ring stage, one loop iteration of a consumer warp) applies up to four edges:
//     A  = / *=  T_pre0[state]          (n_pre >= 1; "=" when the step starts a new partial)
//     A *= T_pre1[state]                (n_pre == 2)
//     A  = P·A  [∘ pop()]               (mv: 1 = set, 2 = multiply with the popped sibling partial)
//     A *= T_post[state]                (n_post == 1, only together with an MV)
// then optionally rescales, pushes A, or stores it as a root-child partial. In a binary tree nearly every step
// carries one MV (30·M DMMAs per warp), so the per-iteration bookkeeping and the tip gathers overlap an MMA burst.
enum : int32_t {
    ST_MV_MASK = 3,          // bits 0-1
    ST_NPRE_SHIFT = 2,       // bits 2-3
    ST_PRE0_SET = 1 << 4,
    ST_NPOST = 1 << 5,
    STF_RESCALE = 1 << 6,
    STF_PUSH = 1 << 7,
    STF_STORE_A = 1 << 8,
    STF_STORE_B = 1 << 9,
    STF_ANY = STF_RESCALE | STF_PUSH | STF_STORE_A | STF_STORE_B,
    ST_PSEUDO_SHIFT = 10,    // bits 10-11: 1-based position (among the step's gathers) of its first clade-table gather, 0 = none
    ST_PSEUDO2_SHIFT = 12,   // bits 12-13: position of a second one (0 = none)
    ST_PSEUDO_TRI = 1 << 14, // the first / second clade gather is a pitchfork (three state rows, 9261 table rows)
    ST_PSEUDO2_TRI = 1 << 15
};
// {word, row of gather 0, 1, 2}, {second rows of the two clade gathers, their third rows}, {first table row of each, 0, 0}
constexpr int STEP_INTS = 12;

struct PruneParams {
    const int4* steps;         // [n_steps][3], see STEP_INTS
    const double* ctab;        // [K][ctab rows][CHERRY_ROW_DOUBLES] clade tables (cherries: 441 rows, pitchforks: 9261)
    int64_t ctab_rate_stride;  // doubles per rate class
    const int32_t* step_off;   // [n_steps] offset (doubles) of the step's matrices inside one rate class's stream
    const double* pstream;     // [K][rate_stride]
    const uint8_t* states;     // device alignment block, states[row*pitch + col]
    int64_t rate_stride;
    int64_t pitch;
    int64_t col0;              // first column of this launch inside the block (multiple of 16)
    double* rootpart;          // [2][K][ncp][20]
    int32_t* rootexp;          // [K][ncp]
    double* spill;             // global overflow of the partial stack
    int64_t ncp;               // padded column count of this launch (n_tiles * S)
    int n_steps;
    int n_tiles;
    int K;
    int smem_levels;           // stack levels resident in shared memory
    int spill_levels;
    int debug;                 // bit 0: proxy fence before the stage release; bit 1: no early barrier probe (diagnostics)
    unsigned* timeline;        // SR_TIMELINE builds only (tools/timeline.py): [warp][SR_TL_STEPS][8] clock stamps of CTA 0
};

#ifdef SR_TIMELINE
constexpr int SR_TL_STEPS = 1024;
#define SR_TL(k) do { if (blockIdx.x == 0 && lane == 0 && it < (uint32_t)SR_TL_STEPS) p.timeline[((size_t)warp * SR_TL_STEPS + it) * 8 + (k)] = (unsigned)clock(); } while (0)
#else
#define SR_TL(k) do { } while (0)
#endif

template <int M, int W_>
struct PruneCfg {
    static constexpr int W = W_;                        // consumer warps (+1 producer warp)
    static constexpr int THREADS = (W_ + 1) * 32;
    static constexpr int S = W * M * 8;                 // columns per work item
    static constexpr int LEVEL_DOUBLES = W * M * 5 * 32;
    static constexpr int LEVEL_BYTES = LEVEL_DOUBLES * 8;
    static constexpr int STATES_OFF = FRAG_BYTES + 3 * TABLE_BYTES;                 // 13920
    static constexpr int STAGE_BYTES = ((STATES_OFF + 7 * S + 127) / 128) * 128;    // 3 gather rows + second and third rows of up to two clade gathers
};

__host__ __device__ constexpr int prune_bar_bytes(int nstage) { return ((2 * nstage * 8 + 127) / 128) * 128; }

// Named-barrier hand-off between the consumer warps of one SM sub-partition (warp id mod 4): a warp may start its MMA
// burst only when the previous warp in the rotation has issued its last DMMA. Without it, two warps that burst at the
// same time interleave, finish together, do their non-MMA work together and leave the tensor pipe idle together.
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// RA: register re-allocation (setmaxnreg). A sub-partition's register file holds three warps of 168 registers, so a
// producer warp next to W_ = 11 consumers leaves its sub-partition one MMA warp short (tensor pipe ~53 % busy there
// against ~79 % on the other three). With RA the CTA is launched as whole warpgroups and the last one (the producer plus
// three warps that exit at once) shrinks to 32 registers, which frees exactly enough for the consumers to grow:
//   M = 2: 16 warps x 128 at launch -> 12 consumers x 160 (three MMA warps on every sub-partition, producer as a light fourth)
//   M = 4: 12 warps x 168 at launch ->  8 consumers x 232 (two MMA warps per sub-partition, four column tiles each)
// burst holes: bit s of the mask = pause (just-in-time fragment fetch) before k-step s
__host__ __device__ constexpr int hole_mask(int holes) {
    if (holes == 0) return 0;
    if (holes == 1) return 1 << 3;
    if (holes == 2) return (1 << 2) | (1 << 4);
    int m = 0;
    while (holes > 0) { m |= 1 << (holes % 10); holes /= 10; }
    return m & 0x1e;
}
__host__ __device__ constexpr int next_hole(int mask, int s) {          // first hole after k-step s, 5 if none
    for (int t = s + 1; t < 5; ++t) if ((mask >> t) & 1) return t;
    return 5;
}
__host__ __device__ constexpr int first_hole(int mask) { return next_hole(mask, 0); }
__host__ __device__ constexpr int pairs_before(int kstep) { return kstep >= 5 ? 8 : (3 * kstep + 1) / 2; }   // 16-byte fragment pairs covering k-steps < kstep

constexpr int prune_threads(int w, bool ra) { return ra ? (w + 4) * 32 : (w + 1) * 32; }

// TWOPS: steps may carry two cherry gathers (schedules of trees with many pairs of sibling cherries).
template <int M, int NSTAGE, int W_, int HOLES = 0, bool ROT = false, bool RA = false, bool TWOPS = false>
__global__ void __launch_bounds__(prune_threads(W_, RA), 1) prune_kernel(PruneParams p) {
    static_assert(!RA || ((W_ == 12 && M == 2) || (W_ == 8 && M == 4)), "register re-allocation is laid out for 12 x M=2 or 8 x M=4 consumer warps");
    using Cfg = PruneCfg<M, W_>;
    constexpr int W = Cfg::W;
    constexpr int S = Cfg::S;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int TILE_DOUBLES = M * 5 * 32;            // one warp's partial on the stack
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stages = smem;                                             // NSTAGE * STAGE_BYTES
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty = full + NSTAGE;
    double* stack = reinterpret_cast<double*>(smem + NSTAGE * STAGE_BYTES + prune_bar_bytes(NSTAGE));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], W);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_items = p.n_tiles * p.K;
    uint32_t it = 0;                                  // ring position, runs on across work items

    // issue the TMA copies of one step (global ring position itp; work item `item_p`, step `step_p`) — one thread
    auto issue_step = [&](uint32_t itp, int item_p, const int4 d, const int4 row_x, const int off_d, bool wait_empty) {
        const int kp = item_p / p.n_tiles, tilep = item_p - kp * p.n_tiles;
        const double* ps = p.pstream + (size_t)kp * p.rate_stride;
        const uint8_t* st = p.states + p.col0 + (int64_t)tilep * S;
        const int stage = itp % NSTAGE;
        const uint32_t ph = (itp / NSTAGE) & 1;
        const int n_leaf = ((d.x >> ST_NPRE_SHIFT) & 3) + ((d.x & ST_NPOST) ? 1 : 0);
        // a cherry gather brings a second state row and no table
        const int pseudo = (((d.x >> ST_PSEUDO_SHIFT) & 3) ? 1 : 0) + (((d.x >> ST_PSEUDO2_SHIFT) & 3) ? 1 : 0);
        const int tri1 = (d.x & ST_PSEUDO_TRI) ? 1 : 0, tri2 = (d.x & ST_PSEUDO2_TRI) ? 1 : 0;
        const uint32_t pbytes = ((d.x & ST_MV_MASK) ? FRAG_BYTES : 0) + (n_leaf - pseudo) * TABLE_BYTES;
        if (wait_empty) mbar_wait(&empty[stage], ph ^ 1);
        unsigned char* dst = stages + stage * STAGE_BYTES;
        mbar_expect_tx(&full[stage], pbytes + (n_leaf + pseudo + tri1 + tri2) * S);
        if (pseudo > 0) bulk_g2s(dst + Cfg::STATES_OFF + 3 * S, st + (int64_t)row_x.x * p.pitch, S, &full[stage]);
        if (pseudo > 1) bulk_g2s(dst + Cfg::STATES_OFF + 4 * S, st + (int64_t)row_x.y * p.pitch, S, &full[stage]);
        if (tri1) bulk_g2s(dst + Cfg::STATES_OFF + 5 * S, st + (int64_t)row_x.z * p.pitch, S, &full[stage]);
        if (tri2) bulk_g2s(dst + Cfg::STATES_OFF + 6 * S, st + (int64_t)row_x.w * p.pitch, S, &full[stage]);
        if (pbytes) bulk_g2s(dst, ps + off_d, pbytes, &full[stage]);
        if (n_leaf > 0) bulk_g2s(dst + Cfg::STATES_OFF, st + (int64_t)d.y * p.pitch, S, &full[stage]);
        if (n_leaf > 1) bulk_g2s(dst + Cfg::STATES_OFF + S, st + (int64_t)d.z * p.pitch, S, &full[stage]);
        if (n_leaf > 2) bulk_g2s(dst + Cfg::STATES_OFF + 2 * S, st + (int64_t)d.w * p.pitch, S, &full[stage]);
    };

    // ===== producer: one lane streams every step's matrices (and tip states) into the ring =====
    auto produce = [&]() {
        if (lane == 0) {
            for (int item = blockIdx.x; item < n_items; item += gridDim.x)
                for (int step = 0; step < p.n_steps; ++step, ++it)
                    issue_step(it, item, __ldg(p.steps + 3 * step), __ldg(p.steps + 3 * step + 1), __ldg(p.step_off + step), true);
        }
    };
    if (RA) {
        // one structural split, so that the register allocator sees which code runs under which budget
        if (warp >= W) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
            if (warp == W) produce();                 // warps 13..15 only existed to donate their registers
            return;
        }
        if (W_ == 12) asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    } else if (warp == W) {
        produce();
        return;
    }

    // ===== consumers: W warps, each walks the whole schedule for its own 8*M columns =====
    const int g = lane >> 2, q = lane & 3;
    double* spill_base = p.spill + ((size_t)blockIdx.x * W + warp) * (size_t)p.spill_levels * TILE_DOUBLES;
    double* my_stack = stack + warp * TILE_DOUBLES + lane;
    const int st_lane_off = Cfg::STATES_OFF + warp * (M * 8) + g;
    // rotation bookkeeping: sub-partition sp4 = warp & 3 holds warps sp4, sp4+4, ...; position rpos among n_rot of them
    const int sp4 = warp & 3, rpos = warp >> 2;
    const int n_rot = (W - sp4 + 3) / 4;
    // ids 0..15: barrier 0 is only used by the one __syncthreads() above, before any hand-off
    const int bar_wait_id = sp4 * 4 + (rpos + n_rot - 1) % n_rot;         // armed by the previous warp in the rotation
    const int bar_pass_id = sp4 * 4 + rpos;
    bool first_burst = (rpos == 0);
    // release of the stage used by ring position `it` (all of this warp's reads of it are complete)
    auto release_stage = [&](int stage) {
        if (lane == 0) mbar_arrive(&empty[stage]);
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int k = item / p.n_tiles, tile = item - k * p.n_tiles;
        const double* ctab_k = p.ctab + (size_t)k * p.ctab_rate_stride + q * 6;
        double A[M][5];
        int cnt[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
            cnt[j] = 0;
#pragma unroll
            for (int r = 0; r < 5; ++r) A[j][r] = 0.0;
        }
        int sp = 0;
        int w_next = __ldg(&p.steps[0].x);
        int2 ro_next = __ldg(reinterpret_cast<const int2*>(p.steps + 2));     // first table rows of the step's clade gathers
        bool next_ready = false;           // result of an early, non-blocking probe of a later stage's full barrier
        // HOLES encodes before which k-steps (1..4) the burst pauses: 0 none, 1 -> {3}, 2 -> {2,4}, digits "ab.." -> {a,b,..}
        constexpr int HMASK = hole_mask(HOLES);
        constexpr int FIRST = pairs_before(next_hole(HMASK, 0));           // fragment pairs fetched before the burst
        int s0[M], s1[M], s2[M];
        double Bf[16];
        for (int step = 0; step < p.n_steps; ++step, ++it) {
            const int stage = it % NSTAGE;
            const uint32_t ph = (it / NSTAGE) & 1;
            const int w = w_next;
            const bool has_next = step + 1 < p.n_steps;
            const int2 ro = ro_next;
            if (has_next) { w_next = __ldg(&p.steps[3 * (step + 1)].x); ro_next = __ldg(reinterpret_cast<const int2*>(p.steps + 3 * (step + 1) + 2)); }
            SR_TL(0);                                   // step begins
            const int mv = w & ST_MV_MASK, n_pre = (w >> ST_NPRE_SHIFT) & 3;
            const unsigned char* stg = stages + stage * STAGE_BYTES;
            const double2* frag = reinterpret_cast<const double2*>(stg) + lane;
            const double* tab = reinterpret_cast<const double*>(stg + (mv ? FRAG_BYTES : 0)) + q * 5;
            const unsigned char* stb = stg + st_lane_off;
            if (!next_ready) mbar_wait(&full[stage], ph);
            next_ready = false;
            SR_TL(1);                                   // stage has landed

            // ---- all shared-memory reads of the step are issued up front
#pragma unroll
            for (int j = 0; j < M; ++j) { s0[j] = stb[j * 8]; s1[j] = stb[S + j * 8]; s2[j] = stb[2 * S + j * 8]; }
            // The burst is issued in HOLES+1 chunks; the fragments of chunk c+1 are loaded only after chunk c was issued,
            // so the warp pauses for one LDS latency inside its own burst. Sibling warps' Hadamard DMULs (same FP64 pipe)
            // slip into those holes instead of starving until the whole burst is over (which bunches the warps).
            if (mv) {
#pragma unroll
                for (int i = 0; i < FIRST; ++i) { const double2 v = frag[i * 32]; Bf[2 * i] = v.x; Bf[2 * i + 1] = v.y; }
            }
            constexpr bool LAZY = (M >= 4);      // big tiles: fetch the post-MMA operands after the burst (registers)
            double pv[M][5];
            if (mv == 2) --sp;
            if (mv == 2 && !LAZY) {
                if (sp < p.smem_levels) {                                  // warp-uniform: LDS, not a generic load
                    const double* src = my_stack + (size_t)sp * (W * TILE_DOUBLES);
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) pv[j][r] = src[(j * 5 + r) * 32];
                } else {
                    const double* src = spill_base + (size_t)(sp - p.smem_levels) * TILE_DOUBLES + lane;
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) pv[j][r] = __ldcg(src + (j * 5 + r) * 32);
                }
            }
            // ---- tips below this node (gathers of P^T rows by observed state): loads first ...
            // At most one of the step's gathers is a collapsed cherry (position ps, 1-based): its row, selected by the state
            // PAIR of the cherry's two tips, comes from the cherry's table in global memory (L2) instead of the stage.
            const int ps = (w >> ST_PSEUDO_SHIFT) & 3, ps2 = TWOPS ? ((w >> ST_PSEUDO2_SHIFT) & 3) : 0;
            const double* crow[M];
            const double* crow2[TWOPS ? M : 1];
            if (ps) {
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    const int sa = ps == 1 ? s0[j] : (ps == 2 ? s1[j] : s2[j]);
                    int ix = sa * 21 + stb[3 * S + j * 8];
                    if (w & ST_PSEUDO_TRI) ix = ix * 21 + stb[5 * S + j * 8];
                    crow[j] = ctab_k + ((size_t)ro.x + ix) * CHERRY_ROW_DOUBLES;
                }
                if (TWOPS && ps2) {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        const int sa = ps2 == 2 ? s1[j] : s2[j];
                        int ix = sa * 21 + stb[4 * S + j * 8];
                        if (w & ST_PSEUDO2_TRI) ix = ix * 21 + stb[6 * S + j * 8];
                        crow2[TWOPS ? j : 0] = ctab_k + ((size_t)ro.y + ix) * CHERRY_ROW_DOUBLES;
                    }
                }
            }
            auto gload = [&](double (&dst)[M][5], const double* const* rows) {
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    double pad;
                    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(dst[j][0]), "=d"(dst[j][1]) : "l"(rows[j]));
                    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2+16];" : "=d"(dst[j][2]), "=d"(dst[j][3]) : "l"(rows[j]));
                    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2+32];" : "=d"(dst[j][4]), "=d"(pad) : "l"(rows[j]));
                }
            };
            // gather number `pos` (1-based) of the step; `tix` = index of its table among the step's staged tables
            auto gather = [&](double (&dst)[M][5], const int (&sx)[M], int pos, int tix) {
                if (ps == pos) {
                    gload(dst, crow);
                } else if (TWOPS && ps2 == pos) {
                    gload(dst, crow2);
                } else {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        const double* row = tab + tix * TABLE_DOUBLES + sx[j] * NS;
#pragma unroll
                        for (int r = 0; r < 5; ++r) dst[j][r] = row[r];
                    }
                }
            };
            double tp0[M][5], tp1[M][5];
            if (n_pre >= 1) gather(tp0, s0, 1, 0);
            if (n_pre == 2) gather(tp1, s1, 2, ps == 1 ? 0 : 1);                 // (staged tables skip the cherry gathers)
            int spost[M];
#pragma unroll
            for (int j = 0; j < M; ++j) spost[j] = n_pre == 0 ? s0[j] : (n_pre == 1 ? s1[j] : s2[j]);
            const int post_tix = n_pre - ((ps && ps <= n_pre) ? 1 : 0) - ((ps2 && ps2 <= n_pre) ? 1 : 0);
            double vp[M][5];
            if (mv && (w & ST_NPOST) && !LAZY) gather(vp, spost, n_pre + 1, post_tix);
            // The post-MMA operands are only consumed after the burst. Pin them here so the compiler cannot sink their loads
            // below the burst (it may move plain loads across the volatile DMMAs): the stage is released right after the
            // burst, and a load from it that is still in flight then races the TMA refill (seen as 1-in-10 runs with a few
            // columns off by 1e-5 at 5,000 taxa; tests/test_gpu_parity.py::*full_size*, tools/stress_big.py).
            if (!LAZY && mv) {
                if ((w & ST_NPOST) && ps != n_pre + 1 && ps2 != n_pre + 1) {   // (a cherry row comes from global memory: it may stay in flight)
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) asm volatile("" : "+d"(vp[j][r]));
                }
                if (mv == 2) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) asm volatile("" : "+d"(pv[j][r]));
                }
            }
            // ... products after all of them are in flight
            if (n_pre == 2 && (w & ST_PRE0_SET)) {                       // cherry: the common case, no selects
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int r = 0; r < 5; ++r) A[j][r] = tp0[j][r] * tp1[j][r];
            } else if (n_pre >= 1) {
                if (w & ST_PRE0_SET) {
                    // (the empty asm keeps this a real branch: if-converted, "A = tp0, then maybe A = tp0 * A_old" makes
                    // the old and the new A live at once, and every step then starts with 20 register moves)
                    asm volatile("");
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) A[j][r] = tp0[j][r];
                } else {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) A[j][r] *= tp0[j][r];
                }
                if (n_pre == 2) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) A[j][r] *= tp1[j][r];
                }
            }
            double acc[M][3][2];
            if (mv) {
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int t = 0; t < 3; ++t) { acc[j][t][0] = 0.0; acc[j][t][1] = 0.0; }
                SR_TL(2);                               // operands requested, ready to burst
                if (ROT && n_rot > 1) {          // rotation: burst only after the previous warp of this sub-partition issued its last DMMA
                    if (!first_burst) bar_sync64(bar_wait_id);
                    first_burst = false;
                }
                SR_TL(3);                               // burst starts
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    // just-in-time fragment fetch: the pairs the next chunk needs are loaded only now, so the warp pauses
                    if ((HMASK >> s) & 1) {
                        asm volatile("" ::: "memory");
                        const int lo = pairs_before(s), hi = pairs_before(next_hole(HMASK, s));
#pragma unroll
                        for (int i = lo; i < hi; ++i) { const double2 v = frag[i * 32]; Bf[2 * i] = v.x; Bf[2 * i + 1] = v.y; }
                    }
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int t = 0; t < 3; ++t) dmma(acc[j][t][0], acc[j][t][1], A[j][s], Bf[s * 3 + t]);
                }
                if (ROT && n_rot > 1) bar_arrive64(bar_pass_id);
                SR_TL(4);                               // last DMMA issued
                if (!LAZY) {
                    // Every read of this stage was issued before or inside the burst (>= 480 cycles ago; the pins above keep
                    // the compiler from sinking them). Shared loads of a warp complete in order, so one instruction that reads
                    // the registers of the LAST loads makes the scoreboard prove all of them complete before the arrive is
                    // issued — the same contract CUTLASS's consumer_release relies on. `debug` bit 0 adds the (5 % slower)
                    // fence.proxy.async for diagnosis.
                    {
                        unsigned sink;
                        asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %1; mov.b64 {%0, hi}, %2; xor.b32 %0, %0, lo; }"
                                     : "=r"(sink) : "d"(vp[M - 1][4]), "d"(pv[M - 1][4]));
                        asm volatile("" :: "r"(sink));
                    }
                    if (p.debug & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    release_stage(stage);
                }
                {   // probe the next stage now: the ~60-cycle TRYWAIT latency hides behind this step's epilogue
                    const uint32_t itp = it + 1;
                    next_ready = (p.debug & 2) ? false : mbar_test_wait(&full[itp % NSTAGE], (itp / NSTAGE) & 1);
                }
                if (LAZY) {
                    if (w & ST_NPOST) gather(vp, spost, n_pre + 1, post_tix);
                    if (mv == 2) {
                        if (sp < p.smem_levels) {
                            const double* src = my_stack + (size_t)sp * (W * TILE_DOUBLES);
#pragma unroll
                            for (int j = 0; j < M; ++j)
#pragma unroll
                                for (int r = 0; r < 5; ++r) pv[j][r] = src[(j * 5 + r) * 32];
                        } else {
                            const double* src = spill_base + (size_t)(sp - p.smem_levels) * TILE_DOUBLES + lane;
#pragma unroll
                            for (int j = 0; j < M; ++j)
#pragma unroll
                                for (int r = 0; r < 5; ++r) pv[j][r] = __ldcg(src + (j * 5 + r) * 32);
                        }
                    }
                }
                SR_TL(5);                               // stage released, next stage probed
                // epilogue: the Hadamard products write A (the registers the next burst reads its operand from) straight
                // from the accumulators. A plain MV (no sibling, no tip) has no product to do that move for free: its result
                // stays in the accumulator registers and the step's tail below works on those (see `finish`).
                if (mv == 2 && (w & ST_NPOST)) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) A[j][r] = (acc[j][r >> 1][r & 1] * pv[j][r]) * vp[j][r];
                } else if (mv == 2) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) A[j][r] = acc[j][r >> 1][r & 1] * pv[j][r];
                } else if (w & ST_NPOST) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) A[j][r] = acc[j][r >> 1][r & 1] * vp[j][r];
                }
                if (LAZY) {
                    // the late gathers are consumed by the products above; pin those in front of the release so that the
                    // (in-order) arrive cannot be issued while a load from this stage is still in flight
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) asm volatile("" : "+d"(A[j][r]));
                    stage_release(&empty[stage], lane);
                }
            } else {
                stage_release(&empty[stage], lane);
            }
            SR_TL(6);                                   // epilogue products issued
            // rescale / push / store of the step's result, wherever it lives: in A, or (plain MV) still in the accumulators.
            // Two instantiations instead of a copy: 20 register moves per step were 6 % of the instructions a warp issues,
            // on an issue port the MMAs already saturate.
            auto finish = [&](auto&& at) {          // at(j, r): reference to entry r of column tile j
            if (w & STF_ANY) {
                if (w & STF_RESCALE) {
                    // Power-of-two rescale done entirely in the integer pipe (the FP64 pipe is the one the MMAs need):
                    // conditionals are >= 0, so the largest one has the largest high word; scaling by 2^-(e-1023)
                    // is a subtraction in the exponent field. Entries more than 2^-1000 below the maximum flush to 0.
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        int hmx = max(max(max(__double2hiint(at(j, 0)), __double2hiint(at(j, 1))),
                                          max(__double2hiint(at(j, 2)), __double2hiint(at(j, 3)))), __double2hiint(at(j, 4)));
                        hmx = max(hmx, __shfl_xor_sync(0xffffffffu, hmx, 1));
                        hmx = max(hmx, __shfl_xor_sync(0xffffffffu, hmx, 2));
                        const int e = (hmx >> 20) & 0x7ff;
                        const int delta = (e - 1023) << 20;
#pragma unroll
                        for (int r = 0; r < 5; ++r) {
                            const int hi = __double2hiint(at(j, r));
                            const int nh = hi - delta;
                            at(j, r) = (e != 0 && nh >= 0x00100000) ? __hiloint2double(nh, __double2loint(at(j, r))) : (e != 0 ? 0.0 : at(j, r));
                        }
                        if (e != 0) cnt[j] += e - 1023;
                    }
                }
                if (w & STF_PUSH) {
                    if (sp < p.smem_levels) {
                        double* dst = my_stack + (size_t)sp * (W * TILE_DOUBLES);
#pragma unroll
                        for (int j = 0; j < M; ++j)
#pragma unroll
                            for (int r = 0; r < 5; ++r) dst[(j * 5 + r) * 32] = at(j, r);
                    } else {
                        double* dst = spill_base + (size_t)(sp - p.smem_levels) * TILE_DOUBLES + lane;
#pragma unroll
                        for (int j = 0; j < M; ++j)
#pragma unroll
                            for (int r = 0; r < 5; ++r) __stcg(dst + (j * 5 + r) * 32, at(j, r));
                    }
                    ++sp;
                }
                if (w & (STF_STORE_A | STF_STORE_B)) {
                    const int which = (w & STF_STORE_B) ? 1 : 0;
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        const int64_t col = (int64_t)tile * S + warp * (M * 8) + j * 8 + g;
                        double* dst = p.rootpart + (((size_t)which * p.K + k) * p.ncp + col) * NS + q * 5;
#pragma unroll
                        for (int r = 0; r < 5; ++r) dst[r] = at(j, r);
                        if (which == 1 && q == 0) p.rootexp[(size_t)k * p.ncp + col] = cnt[j];
                    }
                }
            }
            };
            if (mv == 1 && !(w & ST_NPOST)) {
                finish([&](int j, int r) -> double& { return acc[j][r >> 1][r & 1]; });
                if (!(w & (STF_PUSH | STF_STORE_A | STF_STORE_B))) {     // the next step continues this partial (unary chains)
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) A[j][r] = acc[j][r >> 1][r & 1];
                }
            } else {
                finish([&](int j, int r) -> double& { return A[j][r]; });
            }
            SR_TL(7);                                   // step ends
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
// SMEMP: the K merged root-branch matrices are staged once per CTA in dynamic shared memory, transposed ([k][b][a]) so
// that lane a's 20 reads per rate class are conflict-free; read straight from global memory they are 20 strided
// 8-byte loads per lane and rate class, and the L1 sector traffic of those was what bounded this kernel.
constexpr int ROOT_SMEM_MAX_K = 16;
template <bool SMEMP>
__global__ void __launch_bounds__(ROOT_WARPS * 32) root_kernel(RootParams p) {
    __shared__ double sB[ROOT_WARPS][NS];
    __shared__ double sOut[ROOT_WARPS][NS * NS];
    extern __shared__ __align__(16) double sP[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double LN2 = 0.693147180559945309417232121458;
    if (SMEMP) {
        for (int e = threadIdx.x; e < p.K * NS * NS; e += blockDim.x) {
            const int k = e / (NS * NS), ab = e - k * (NS * NS), a = ab / NS, b = ab - a * NS;
            sP[k * (NS * NS) + b * NS + a] = p.proot[e];
        }
        __syncthreads();
    }
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
                if (SMEMP) {
                    const double* pr = sP + k * (NS * NS) + lane;
#pragma unroll
                    for (int b = 0; b < NS; ++b) row[b] = fma(fa * pr[b * NS], sB[warp][b], row[b]);
                } else {
                    const double* pr = p.proot + (size_t)k * NS * NS + lane * NS;
#pragma unroll
                    for (int b = 0; b < NS; ++b) row[b] = fma(fa * pr[b], sB[warp][b], row[b]);
                }
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
