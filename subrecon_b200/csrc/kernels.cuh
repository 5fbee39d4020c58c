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
constexpr int TAB_ROW_DOUBLES = 24;               // a tip-table row: 4 groups of {5 values, 1 pad}, like a clade-table row —
                                                  // the lane owning states 5q..5q+4 fetches its group with three 16-byte loads
constexpr int TABLE_DOUBLES = 21 * TAB_ROW_DOUBLES;   // P^T rows for states 0..19 + the missing-data row
constexpr int TABLE_BYTES = TABLE_DOUBLES * 8;    // 4032
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
// The same with an L2 eviction-priority hint (createpolicy): evict_first for bytes that are read once and never again,
// evict_last for bytes every CTA keeps re-reading
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
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
        // JointBranchReconstruction.java:201-205, so its contribution is the row sum, left to right). Entry i of row s is
        // stored at s*24 + (i/5)*6 + i%5; the sixth slot of every group is padding.
        const bool ident = (kind == SLOT_IDENTITY);
        for (int e = tid; e < TABLE_DOUBLES; e += 128) {
            const int s = e / TAB_ROW_DOUBLES, c = e % TAB_ROW_DOUBLES, i = (c / 6) * 5 + c % 6;
            double v;
            if (c % 6 == 5) v = 0.0;
            else if (s < NS) v = ident ? (i == s ? 1.0 : 0.0) : P[i * NS + s];
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
// hash of its state codes; columns with equal hashes are then compared byte for byte (verify_patterns_kernel).
__device__ __forceinline__ uint64_t mix64(uint64_t x) {          // murmur3 finaliser
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256) hash_columns_kernel(const uint8_t* __restrict__ states, int64_t pitch, int n_taxa,
                                                           int64_t n, uint64_t* __restrict__ h1, uint64_t* __restrict__ h2,
                                                           uint32_t* __restrict__ idx, int hashed_taxa) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    uint64_t a = 0x9e3779b97f4a7c15ULL, b = 0xc2b2ae3d27d4eb4fULL;
    // (hashed_taxa < n_taxa only in the test mode that provokes collisions, option "compress_patterns" = 2)
    for (int t = 0; t < hashed_taxa; ++t) {                      // adjacent threads read adjacent bytes of one row
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
// Equal hashes are only a claim: every column is compared byte for byte with the representative of the pattern it was
// mapped to (the gathered pattern matrix is small and stays in L2; the alignment block is read once, coalesced). Any
// difference raises the flag and the caller recomputes the block the plain way.
__global__ void __launch_bounds__(256) verify_patterns_kernel(const uint8_t* __restrict__ states, int64_t pitch, int n_taxa, int64_t n,
                                                              const uint32_t* __restrict__ map, const uint8_t* __restrict__ pat,
                                                              int64_t pat_pitch, int* __restrict__ mismatch) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const uint32_t u = map[c];
    const int t0 = blockIdx.y * 32, t1 = min(n_taxa, t0 + 32);
    bool bad = false;
    for (int t = t0; t < t1; ++t) bad |= states[(int64_t)t * pitch + c] != pat[(int64_t)t * pat_pitch + u];
    if (bad) atomicOr(mismatch, 1);
}
// results of the distinct patterns back to every original column: one warp per column
__global__ void __launch_bounds__(256) scatter_patterns_kernel(const uint32_t* __restrict__ map, int64_t n, int64_t out0,
                                                               const double* __restrict__ u_lnl, const double* __restrict__ u_joint,
                                                               const uint64_t* __restrict__ u_compact, double* __restrict__ lnl,
                                                               double* __restrict__ joint, uint64_t* __restrict__ compact, int compact_words) {
    const int lane = threadIdx.x & 31;
    for (int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const uint32_t u = map[c];
        if (lnl && lane == 0) lnl[out0 + c] = u_lnl[u];
        if (joint) {
            const double2* src = reinterpret_cast<const double2*>(u_joint + (size_t)u * 400);
            double2* dst = reinterpret_cast<double2*>(joint + (size_t)(out0 + c) * 400);
            for (int e = lane; e < 200; e += 32) dst[e] = src[e];
        }
        if (compact)
            for (int e = lane; e < compact_words; e += 32) compact[(size_t)(out0 + c) * compact_words + e] = u_compact[(size_t)u * compact_words + e];
    }
}

// K2 (prune_kernel) and K0b (clade_index_kernel) live in prune.cuh, included at the end of this file.

// ------------------------------------------------------------------------------------------- K3
struct RootParams {
    const double* rootpart;    // [2][K][ncp][20]
    const int32_t* rootexp;    // [K][ncp]
    const double* proot;       // [K][400]
    const double* pi;          // [20]
    double* lnl;               // [n] or null
    double* joint;             // [n][400] or null
    void* compact;             // compact records [n] or null
    int compact_cap;           // entries per record; record stride compact_stride(cap)
    double threshold;
    int64_t ncp;
    int64_t n;                 // output columns of this launch
    int64_t col_skip;          // launch column of output 0 (launches start on a 16-column boundary)
    int64_t out0;              // offset of this launch's first output in the output arrays
    int K;
};

struct CompactRec {            // must match sr_compact in include/subrecon_b200.h (the layout for capacity 8)
    double lnl, max_ii, max_p;
    int32_t n_kept, interesting;
    int16_t pair[8];
    double prob[8];
};
// a record of capacity cap: 32-byte head, int16 pair[cap] padded to 8 bytes, double prob[cap]
__host__ __device__ constexpr int compact_pairs_bytes(int cap) { return (2 * cap + 7) / 8 * 8; }
__host__ __device__ constexpr int compact_stride(int cap) { return 32 + compact_pairs_bytes(cap) + 8 * cap; }

// One warp per ROOT_COLS columns; lane a < 20 owns row a of each column's 20x20 joint matrix.
// joint[a][b] ∝ Σ_k 2^(e_k-e_max) · π_a · A_k[a] · P_k^{tA+tB}[a][b] · B_k[b]: for every (k, b) a lane needs one entry of the
// root-branch matrix (its own row: one conflict-free LDS) and B_k[b] of each column (a broadcast LDS). Shared-memory
// wavefronts are what bound this kernel, so a warp works on ROOT_COLS = 2 columns at once: the matrix entry is loaded once
// for both and the two B_k[b] sit side by side (one LDS.128 broadcast) — half the wavefronts per column. The arithmetic per
// column is unchanged (same operations in the same order as with one column per warp).
constexpr int ROOT_WARPS = 8;
constexpr int ROOT_COLS = 2;
// SMEMP: the K merged root-branch matrices are staged once per CTA in dynamic shared memory, transposed ([k][b][a]) so
// that lane a's 20 reads per rate class are conflict-free; read straight from global memory they are 20 strided
// 8-byte loads per lane and rate class, and the L1 sector traffic of those was what bounded this kernel.
constexpr int ROOT_SMEM_MAX_K = 16;
template <bool SMEMP>
__global__ void __launch_bounds__(ROOT_WARPS * 32, 2) root_kernel(RootParams p) {
    constexpr int C = ROOT_COLS;
    __shared__ __align__(16) double sB[ROOT_WARPS][NS][C];
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
    const double pia = lane < NS ? p.pi[lane] : 0.0;
    for (int64_t grp = (int64_t)blockIdx.x * ROOT_WARPS + warp; grp * C < p.n; grp += (int64_t)gridDim.x * ROOT_WARPS) {
        const int64_t oc0 = grp * C;
        const int ncol = (int)min((int64_t)C, p.n - oc0);
        int64_t col[C];
        int emax[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            col[c] = oc0 + (c < ncol ? c : 0) + p.col_skip;              // a missing second column repeats the first (result unused)
            emax[c] = INT_MIN;
            for (int k = 0; k < p.K; ++k) emax[c] = max(emax[c], p.rootexp[(size_t)k * p.ncp + col[c]]);
        }
        double row[C][NS];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int b = 0; b < NS; ++b) row[c][b] = 0.0;
        for (int k = 0; k < p.K; ++k) {
            double fa[C];
            __syncwarp();
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int e = p.rootexp[(size_t)k * p.ncp + col[c]];
                const double wk = ldexp(1.0, e - emax[c]);                               // exact, <= 1
                const double* pa = p.rootpart + (((size_t)0 * p.K + k) * p.ncp + col[c]) * NS;
                const double* pb = p.rootpart + (((size_t)1 * p.K + k) * p.ncp + col[c]) * NS;
                if (lane < NS) {
                    sB[warp][lane][c] = pb[lane];
                    fa[c] = wk * (pia * pa[lane]);
                }
            }
            __syncwarp();
            if (lane < NS) {
                if (SMEMP) {
                    const double* pr = sP + k * (NS * NS) + lane;
#pragma unroll
                    for (int b = 0; b < NS; ++b) {
                        const double prb = pr[b * NS];
#pragma unroll
                        for (int c = 0; c < C; ++c) row[c][b] = fma(fa[c] * prb, sB[warp][b][c], row[c][b]);
                    }
                } else {
                    const double* pr = p.proot + (size_t)k * NS * NS + lane * NS;
#pragma unroll
                    for (int b = 0; b < NS; ++b) {
                        const double prb = pr[b];
#pragma unroll
                        for (int c = 0; c < C; ++c) row[c][b] = fma(fa[c] * prb, sB[warp][b][c], row[c][b]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (c >= ncol) break;
            const int64_t oc = oc0 + c;
            double rs = 0.0;
#pragma unroll
            for (int b = 0; b < NS; ++b) rs += row[c][b];
            double tot = rs;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            const double inv = 1.0 / tot;
            const double lnl = log(tot) + (double)emax[c] * LN2 - log((double)p.K);
            if (p.lnl && lane == 0) p.lnl[p.out0 + oc] = lnl;
            if (p.joint || p.compact) {
                __syncwarp();
                if (lane < NS) {
#pragma unroll
                    for (int b = 0; b < NS; ++b) sOut[warp][lane * NS + b] = row[c][b] * inv;
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
                    const int cap = p.compact_cap;
                    unsigned char* recb = reinterpret_cast<unsigned char*>(p.compact) + (size_t)(p.out0 + oc) * compact_stride(cap);
                    CompactRec* rec = reinterpret_cast<CompactRec*>(recb);                  // the head is common to every capacity
                    int16_t* pair = reinterpret_cast<int16_t*>(recb + 32);
                    double* prob = reinterpret_cast<double*>(recb + 32 + compact_pairs_bytes(cap));
                    for (int base = 0; base < NS * NS; base += 32) {
                        const int e = base + lane;
                        const bool keep = e < NS * NS && sOut[warp][e] >= p.threshold;
                        const unsigned m = __ballot_sync(0xffffffffu, keep);
                        const int pos = nk + __popc(m & ((1u << lane) - 1));
                        if (keep && pos < cap) { pair[pos] = (int16_t)e; prob[pos] = sOut[warp][e]; }
                        nk += __popc(m);
                    }
                    if (lane == 0) {
                        rec->lnl = lnl; rec->max_ii = mxii; rec->max_p = mxp; rec->n_kept = nk;
                        rec->interesting = (mxii <= 1.0 - p.threshold && mxp >= p.threshold) ? 1 : 0;
                    }
                    for (int e = nk + lane; e < compact_pairs_bytes(cap) / 2; e += 32) pair[e] = -1;      // unused slots (and the padding)
                    for (int e = nk + lane; e < cap; e += 32) prob[e] = 0.0;
                }
                __syncwarp();                                                           // sOut is reused by the next column
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

#include "prune.cuh"
