// prune.cuh — K2, the post-order pruning kernel (recon/JointBranchReconstruction.java:191-247 for a tile of columns).
//
// Persistent CTA, 12 consumer warps x 16 columns + one TMA producer warp on re-allocated registers (setmaxnreg).
// A pruning STEP (one ring stage, one loop iteration of a consumer warp) applies up to four edges:
//     A  = G0 ∘ G1   |   A *= G0   |   A = G0        (PRE: gathers that start or extend the partial)
//     R  = P·A [∘ pop()]                             (MV: 30 DMMA.8x8x4 per warp; 2 = multiply with the popped sibling)
//     R *= Gpost                                     (POST)
// then optionally rescales, pushes or stores R as a root-child partial. A gather G is a row of 20 conditionals selected
// by the column's data: a tip's P^T row by observed state (table staged in shared memory by TMA) or a collapsed clade's
// row by the state combination of its tips (K1b table in global memory, index precombined by K0b).
//
// Everything that can be decided on the host is: the schedule arrives as one 16-byte record per step for the consumers
// (shape id, tail flags, stack levels, ready-made byte offsets of the step's gather sources) and one 32-byte record for
// the producer (bytes to copy, rows to fetch). The consumer loop switches on the shape id into a body specialised at
// compile time for the shapes that make up nearly all of a schedule, and into one general body for the rest, so a
// warp-step issues ~150 instructions around its 30 DMMAs instead of ~350 (the round-1 kernel decoded a bit-field per
// step; its FP64 tensor pipe was 63 % active and the issue slots next to the DMMAs were what bounded it).
#pragma once
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

// L2 residency hints (0 = none, 1 = TMA copies carry eviction priorities, 2 = and the root partials leave with streaming
// stores). Measured at BASELINE config #3 with distinct columns: DRAM reads 10.0 -> 7.4 GB per 1 M-column launch, 69.9 -> 68.5 ms.
#ifndef SR_L2_HINTS
#define SR_L2_HINTS 2
#endif

namespace srk {

// ---- geometry (one layout: the tuning variants of round 1 never won and are gone)
constexpr int P2_M = 2;                               // column tiles (8 columns) per consumer warp
constexpr int P2_W = 12;                              // consumer warps: three per SM sub-partition
constexpr int P2_NSTAGE = 4;
constexpr int P2_S = P2_W * P2_M * 8;                 // 192 columns per work item
constexpr int P2_THREADS = (P2_W + 4) * 32;           // launched as four warpgroups; the fourth donates its registers
constexpr int P2_TAB_OFF = FRAG_BYTES;                // stage: [MV fragments 4096][<= 3 tip tables 4032 each][3 index rows 384 each][record 16]
constexpr int P2_ROWS_OFF = FRAG_BYTES + 3 * TABLE_BYTES;
constexpr int P2_ROW_BYTES = 2 * P2_S;                // a gather's index row: uint8 states (tips) or uint16 clade indices
constexpr int P2_REC_OFF = P2_ROWS_OFF + 3 * P2_ROW_BYTES;   // the step's 16-byte consumer record, copied in with everything else
constexpr int P2_STAGE_BYTES = ((P2_REC_OFF + 16 + 127) / 128) * 128;
constexpr int P2_LANE_BYTES = P2_M * 5 * 8;           // a lane's slice of a stacked partial: 10 doubles, contiguous (5 x LDS.128)
constexpr int P2_WARP_BYTES = 32 * P2_LANE_BYTES;
constexpr int P2_LEVEL_BYTES = P2_W * P2_WARP_BYTES;  // 30720
constexpr int P2_BAR_BYTES = 128;

// ---- consumer record, word x
enum : uint32_t {
    CR_SHAPE_MASK = 7,       // bits 0-2: 0..6 = specialised bodies, 7 = general body
    CR_RESCALE = 1u << 3,
    CR_PUSH = 1u << 4,
    CR_STORE_A = 1u << 5,
    CR_STORE_B = 1u << 6,
    CR_TAIL = CR_RESCALE | CR_PUSH | CR_STORE_A | CR_STORE_B,
    CR_CLADE0 = 1u << 7,     // bits 7-9: gather slot i reads a clade table (global) instead of a staged tip table
    CR_POP_SHIFT = 10,       // bits 10-14: stack level popped by an MV of kind 2
    CR_PUSH_SHIFT = 15,      // bits 15-19: stack level pushed to
    CR_NPRE_SHIFT = 20,      // bits 20-21: general body only
    CR_PRE_SET = 1u << 22,
    CR_MV_SHIFT = 23,        // bits 23-24
    CR_POST = 1u << 25
};
// shape ids: {PRE, MV, POST}
enum : int { SH_MV1 = 0, SH_MV2 = 1, SH_MV1_POST = 2, SH_NEW2_MV1 = 3, SH_NEW2_MV1_POST = 4, SH_NEW2_MV2 = 5, SH_MV2_POST = 6, SH_GENERAL = 7 };
enum : int { PRE_NONE = 0, PRE_NEW2 = 2 };

// ---- the schedule as build_schedule emits it and sr_schedule_dump exports it (STEP_INTS ints per step); build_records
//      turns it into the two record streams above
enum : int32_t {
    ST_MV_MASK = 3,          // bits 0-1: 1 = R = P·A, 2 = R = pop() ∘ P·A
    ST_NPRE_SHIFT = 2,       // bits 2-3: gathers before the MV
    ST_PRE0_SET = 1 << 4,    // the first gather starts a new partial
    ST_NPOST = 1 << 5,       // one gather after the MV
    STF_RESCALE = 1 << 6,
    STF_PUSH = 1 << 7,
    STF_STORE_A = 1 << 8,
    STF_STORE_B = 1 << 9,
    ST_CLADE_SHIFT = 10,     // bits 10-12: gather g of the step reads a collapsed clade's table
    ST_TRI_SHIFT = 13        // bits 13-15: ... and that clade is a pitchfork (three state rows, 9261 table rows)
};
// {word, alignment row of gather 0, 1, 2, second row of clade gather 0, 1, 2, third row of each, first table row of each, 0, 0, 0}
constexpr int STEP_INTS = 16;

struct PruneParams {
    const uint4* crec;         // [n_steps] consumer records {x, offset of gather slot 0, 1, 2}: a tip slot's offset is the byte
                               // offset of its staged table inside the stage, a clade slot's the byte offset of its table
                               // inside one rate class's block of ctab
    const uint4* prec;         // [n_steps][2] producer records {stream byte offset, matrix bytes, n rows | kinds << 2, row 0}, {row 1, row 2, -, -}
    const double* ctab;        // [K][ctab rows][CHERRY_ROW_DOUBLES] clade tables (cherries: 441 rows, pitchforks: 9261)
    int64_t ctab_rate_stride;  // doubles per rate class
    const double* pstream;     // [K][rate_stride]
    int64_t rate_stride;
    const uint8_t* states;     // device alignment block, states[row*pitch + col], codes 0..20
    int64_t pitch;
    int64_t col0;              // first column of this launch inside the block (multiple of 16)
    const uint16_t* cidx;      // [n clade gathers][ncp] precombined table-row indices (K0b)
    double* rootpart;          // [2][K][ncp][20]
    int32_t* rootexp;          // [K][ncp]
    double* spill;             // global overflow of the partial stack
    int64_t ncp;               // padded column count of this launch (n_tiles * S)
    int n_steps;
    int n_tiles;
    int K;
    int smem_levels;           // stack levels resident in shared memory
    int spill_levels;
    uint32_t zero;             // 0, but the compiler cannot know: ties the stage release to the data read from the stage
    int k_fastest;             // work items ordered tile-major (the K rate classes of a tile run concurrently: states are read from HBM once)
    unsigned* timeline;        // SR_TIMELINE builds only
};

#ifdef SR_TIMELINE
constexpr int SR_TL_STEPS = 1024;
#define SR_TL(k) do { if (blockIdx.x == 0 && lane == 0 && it < (uint32_t)SR_TL_STEPS) p.timeline[((size_t)warp * SR_TL_STEPS + it) * 8 + (k)] = (unsigned)clock(); } while (0)
#else
#define SR_TL(k) do { } while (0)
#endif

// ---- shared-memory accessors on 32-bit shared addresses (no generic-pointer arithmetic in the loop)
__device__ __forceinline__ void lds128(double& a, double& b, uint32_t addr) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, double a, double b) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void ldg128(double& a, double& b, const char* ptr) {
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(ptr));
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// per-warp state of the consumer loop that lives across steps
struct ConsumerState {
    double A[P2_M][5];          // the running partial: lane 4g+q holds states 5q..5q+4 of column g of each tile
    int cnt[P2_M];              // power-of-two exponent taken out of A so far
};

__global__ void __launch_bounds__(P2_THREADS, 1) prune_kernel(PruneParams p) {
    constexpr int M = P2_M, W = P2_W, NSTAGE = P2_NSTAGE, S = P2_S;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + NSTAGE * P2_STAGE_BYTES;       // full[s] at full0 + 8 s, empty[s] at full0 + 8 (NSTAGE + s)
    const uint32_t stack0 = full0 + P2_BAR_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * P2_STAGE_BYTES);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&full[NSTAGE + s], W);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_items = p.n_tiles * p.K;

    // ===================================================================== producer: one lane feeds the ring by TMA
    if (warp >= W) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp != W || lane != 0) return;                       // warps 13..15 only existed to donate their registers
        uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * P2_STAGE_BYTES);
        uint32_t stage = 0, phase = 0;
#if SR_L2_HINTS
        // L2 residency by what the bytes are: the index rows (alignment states, precombined clade indices: GBs per launch)
        // are read exactly once — evict_first, so that they do not push the clade tables' hot rows out of L2; the matrix
        // stream and the records are re-read by every CTA for every tile — evict_last.
        const uint64_t pol_once = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
#endif
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int k = p.k_fastest ? item % p.K : item / p.n_tiles;
            const int tile = p.k_fastest ? item / p.K : item - k * p.n_tiles;
            const char* ps = reinterpret_cast<const char*>(p.pstream + (size_t)k * p.rate_stride);
            const uint8_t* st = p.states + p.col0 + (int64_t)tile * S;
            const uint16_t* ci = p.cidx + (int64_t)tile * S;
            for (int step = 0; step < p.n_steps; ++step) {
                const uint4 a = __ldg(p.prec + 2 * step);
                const uint4 b = __ldg(p.prec + 2 * step + 1);
                const uint32_t n_rows = a.z & 3, kinds = a.z >> 2;
                while (!mbar_try_wait(&full[NSTAGE + stage], phase ^ 1)) __nanosleep(64);   // (a spinning probe steals issue slots from sub-partition 0)
                unsigned char* dst = smem + stage * P2_STAGE_BYTES;
                uint32_t row_bytes = 0;
                for (uint32_t r = 0; r < n_rows; ++r) row_bytes += ((kinds >> r) & 1) ? 2 * S : S;
                mbar_expect_tx(&full[stage], a.y + row_bytes + 16);
#if SR_L2_HINTS
                bulk_g2s_hint(dst + P2_REC_OFF, p.crec + step, 16, &full[stage], pol_keep);
                if (a.y) bulk_g2s_hint(dst, ps + a.x, a.y, &full[stage], pol_keep);
                for (uint32_t r = 0; r < n_rows; ++r) {
                    const uint32_t row = r == 0 ? a.w : (r == 1 ? b.x : b.y);
                    if ((kinds >> r) & 1) bulk_g2s_hint(dst + P2_ROWS_OFF + r * P2_ROW_BYTES, ci + (int64_t)row * p.ncp, 2 * S, &full[stage], pol_once);
                    else bulk_g2s_hint(dst + P2_ROWS_OFF + r * P2_ROW_BYTES, st + (int64_t)row * p.pitch, S, &full[stage], pol_once);
                }
#else
                bulk_g2s(dst + P2_REC_OFF, p.crec + step, 16, &full[stage]);
                if (a.y) bulk_g2s(dst, ps + a.x, a.y, &full[stage]);
                for (uint32_t r = 0; r < n_rows; ++r) {
                    const uint32_t row = r == 0 ? a.w : (r == 1 ? b.x : b.y);
                    if ((kinds >> r) & 1) bulk_g2s(dst + P2_ROWS_OFF + r * P2_ROW_BYTES, ci + (int64_t)row * p.ncp, 2 * S, &full[stage]);
                    else bulk_g2s(dst + P2_ROWS_OFF + r * P2_ROW_BYTES, st + (int64_t)row * p.pitch, S, &full[stage]);
                }
#endif
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");

    // ===================================================================== consumers
    const int g = lane >> 2, q = lane & 3;
    const uint32_t col_in_tile = warp * (M * 8) + g;                       // this lane's column (tile j adds 8 j)
    const uint32_t q48 = q * 48;                                           // a lane's group of a 24-double table row
    uint32_t my_stack = stack0 + warp * P2_WARP_BYTES + lane * P2_LANE_BYTES;
    uint32_t rows_lane = P2_ROWS_OFF + col_in_tile;                       // byte offset of this lane's state byte inside a stage
    uint32_t frag_lane = lane * 16;
    uint32_t rows2_lane = P2_ROWS_OFF + 2 * col_in_tile;                  // ... of its 16-bit clade index
    uint32_t q48r = q48, lane_r = lane;
    uint32_t zero = p.zero;
    // (opaque to the optimiser: otherwise each of these is re-derived from %tid in every step)
    asm volatile("" : "+r"(my_stack), "+r"(rows_lane), "+r"(rows2_lane), "+r"(frag_lane), "+r"(q48r), "+r"(lane_r), "+r"(zero));
    // stack levels that did not fit in shared memory live in an L2-resident global buffer; the address is formed where it
    // is needed (deep trees only) instead of occupying two registers across the whole loop
    auto spill_ptr = [&](int lvl) {
        return reinterpret_cast<char*>(p.spill) + (((size_t)blockIdx.x * W + warp) * (size_t)p.spill_levels + (size_t)(lvl - p.smem_levels)) * P2_WARP_BYTES + lane * P2_LANE_BYTES;
    };
    // Burst rotation among the three warps of an SM sub-partition (warp & 3): a warp starts its 30 DMMAs only when the
    // previous warp in the rotation has issued its last one, so bursts run back to back instead of interleaving. The last
    // warp of each rotation arms the first one's barrier once, up front.
    const int sp4 = warp & 3, rpos = warp >> 2;
    constexpr int N_ROT = W / 4;
    const int bar_wait_id = sp4 * 4 + (rpos + N_ROT - 1) % N_ROT + 1;      // ids 1..15 (0 is __syncthreads')
    const int bar_pass_id = sp4 * 4 + rpos + 1;
    if (rpos == N_ROT - 1) bar_arrive64(bar_pass_id);

    uint32_t stage_addr = smem0, full_addr = full0, phase = 0, ring = 0;     // ring position: stage address, its barrier, its parity
#ifdef SR_TIMELINE
    uint32_t it = 0;
#endif
    bool next_ready = false;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int k = p.k_fastest ? item % p.K : item / p.n_tiles;
        const int tile = p.k_fastest ? item / p.K : item - k * p.n_tiles;
        const char* const ctab_kq = reinterpret_cast<const char*>(p.ctab + (size_t)k * p.ctab_rate_stride) + q48;
        ConsumerState cs;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            cs.cnt[j] = 0;
#pragma unroll
            for (int r = 0; r < 5; ++r) cs.A[j][r] = 0.0;
        }

        for (int step = 0; step < p.n_steps; ++step) {
            SR_TL(0);
            if (!next_ready) mbar_wait_u32(full_addr, phase);
            next_ready = false;
            SR_TL(1);
            const uint32_t stg = stage_addr;
            const uint4 rec = lds_u32x4(stg + P2_REC_OFF);                // arrived with the step's matrices
            const uint32_t rows = stg + rows_lane;                         // tip slots: + slot*384 + 8 j ; clade slots: + col_in_tile again

            // ---- building blocks -------------------------------------------------------------------------------------
            // A gather fills d with the 5 conditionals of this lane's states for each of its M columns, in two phases so that
            // the index loads of all of a step's gathers are in flight before the first dependent row load is issued.
            auto gather_idx = [&](uint32_t (&ix)[M], int slot) {
                if (rec.x & (CR_CLADE0 << slot)) {
#pragma unroll
                    for (int j = 0; j < M; ++j) ix[j] = lds_u16(stg + rows2_lane + slot * P2_ROW_BYTES + 16 * j);
                } else {
#pragma unroll
                    for (int j = 0; j < M; ++j) ix[j] = lds_u8(rows + slot * P2_ROW_BYTES + 8 * j);
                }
            };
            auto gather_rows = [&](double (&d)[M][5], const uint32_t (&ix)[M], int slot, uint32_t off) {
                if (rec.x & (CR_CLADE0 << slot)) {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        const char* ptr = ctab_kq + off + (size_t)ix[j] * (CHERRY_ROW_DOUBLES * 8);
                        double pad;
                        ldg128(d[j][0], d[j][1], ptr);
                        ldg128(d[j][2], d[j][3], ptr + 16);
                        ldg128(d[j][4], pad, ptr + 32);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        const uint32_t a = stg + off + q48r + ix[j] * (TAB_ROW_DOUBLES * 8);
                        double pad;
                        lds128(d[j][0], d[j][1], a);
                        lds128(d[j][2], d[j][3], a + 16);
                        lds128(d[j][4], pad, a + 32);
                    }
                }
            };
            // Release of the stage to the producer. An mbarrier arrive does not wait for this warp's outstanding ld.shared, and
            // ptxas schedules the arrive freely among register-only instructions (it hoisted it above the burst's last DMMAs,
            // whose fragment loads were then still in flight when the TMA refill began: rare wrong matrix entries in the warps
            // that release last). So the arrive's ADDRESS is made to depend on registers that every load from the stage feeds:
            // `dep` is OR-ed from such values and masked with a kernel parameter that is zero.
            auto release = [&](int dep) {
                const uint32_t z = (uint32_t)dep & zero;
                if (lane_r == 0) mbar_arrive_u32(full_addr + 8 * NSTAGE + z);
            };
            auto dep_of_A = [&]() {        // every value of the partial: covers a gather's three row loads per column tile
                int d = 0;
#pragma unroll
                for (int j = 0; j < M; ++j) d |= __double2hiint(cs.A[j][1]) | __double2hiint(cs.A[j][3]) | __double2hiint(cs.A[j][4]);
                return d;
            };
            auto pop = [&](double (&d)[M][5]) {
                const int lvl = (rec.x >> CR_POP_SHIFT) & 31;
                if (lvl < p.smem_levels) {
                    const uint32_t a = my_stack + lvl * P2_LEVEL_BYTES;
#pragma unroll
                    for (int i = 0; i < 5; ++i) lds128(d[(2 * i) / 5][(2 * i) % 5], d[(2 * i + 1) / 5][(2 * i + 1) % 5], a + 16 * i);
                } else {
                    const double2* src = reinterpret_cast<const double2*>(spill_ptr(lvl));
#pragma unroll
                    for (int i = 0; i < 5; ++i) { const double2 v = __ldcg(src + i); d[(2 * i) / 5][(2 * i) % 5] = v.x; d[(2 * i + 1) / 5][(2 * i + 1) % 5] = v.y; }
                }
            };
            // rescale / push / store of the step's result
            auto tail = [&]() {
                auto at = [&](int j, int r) -> double& { return cs.A[j][r]; };
                if (rec.x & CR_RESCALE) {
                    // Power-of-two rescale on the integer pipe (the FP64 pipe belongs to the MMAs): conditionals are >= 0, so
                    // the largest has the largest high word; dividing by 2^(e-1023) is a subtraction in the exponent field.
                    // An entry more than 2^-1022 below the column's maximum ends as a subnormal (~0 next to a maximum of 1..2).
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        int hmx = max(max(max(__double2hiint(at(j, 0)), __double2hiint(at(j, 1))),
                                          max(__double2hiint(at(j, 2)), __double2hiint(at(j, 3)))), __double2hiint(at(j, 4)));
                        hmx = max(hmx, __shfl_xor_sync(0xffffffffu, hmx, 1));
                        hmx = max(hmx, __shfl_xor_sync(0xffffffffu, hmx, 2));
                        const int e = hmx >> 20;                                   // sign bit is 0
                        const int delta = e ? ((e - 1023) << 20) : 0;              // an all-zero / subnormal column is left alone
#pragma unroll
                        for (int r = 0; r < 5; ++r) {
                            const int nh = max(__double2hiint(at(j, r)) - delta, 0);
                            at(j, r) = __hiloint2double(nh, __double2loint(at(j, r)));
                        }
                        cs.cnt[j] += delta >> 20;
                    }
                }
                if (rec.x & CR_PUSH) {
                    const int lvl = (rec.x >> CR_PUSH_SHIFT) & 31;
                    if (lvl < p.smem_levels) {
                        const uint32_t a = my_stack + lvl * P2_LEVEL_BYTES;
#pragma unroll
                        for (int i = 0; i < 5; ++i) sts128(a + 16 * i, at((2 * i) / 5, (2 * i) % 5), at((2 * i + 1) / 5, (2 * i + 1) % 5));
                    } else {
                        double2* dst = reinterpret_cast<double2*>(spill_ptr(lvl));
#pragma unroll
                        for (int i = 0; i < 5; ++i) __stcg(dst + i, make_double2(at((2 * i) / 5, (2 * i) % 5), at((2 * i + 1) / 5, (2 * i + 1) % 5)));
                    }
                }
                if (rec.x & (CR_STORE_A | CR_STORE_B)) {
                    const int which = (rec.x & CR_STORE_B) ? 1 : 0;
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        const int64_t col = (int64_t)tile * S + col_in_tile + j * 8;
                        double* dst = p.rootpart + (((size_t)which * p.K + k) * p.ncp + col) * NS + q * 5;
#pragma unroll
#if SR_L2_HINTS >= 2
                        for (int r = 0; r < 5; ++r) __stcs(dst + r, at(j, r));        // written once, read once by K3: streaming stores
                        if (which == 1 && q == 0) __stcs(p.rootexp + (size_t)k * p.ncp + col, cs.cnt[j]);
#else
                        for (int r = 0; r < 5; ++r) dst[r] = at(j, r);
                        if (which == 1 && q == 0) p.rootexp[(size_t)k * p.ncp + col] = cs.cnt[j];
#endif
                    }
                }
            };
            // the 30-DMMA burst of one contraction, handed over from / to the neighbouring warps of the sub-partition
            auto burst = [&](double (&acc)[M][3][2], double (&Bf)[16], uint32_t frag) {
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int t = 0; t < 3; ++t) { acc[j][t][0] = 0.0; acc[j][t][1] = 0.0; }
                SR_TL(2);
                bar_sync64(bar_wait_id);
                SR_TL(3);
                // Issued in three chunks (k-steps {0,1}, {2,3}, {4}); the fragments of the next chunk are loaded only after the
                // previous one was issued, so the warp pauses for one LDS latency inside its own burst and the sibling warps'
                // Hadamard DMULs (same FP64 pipe, starved while DMMAs stream) slip into those holes.
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    if (s == 2) {
#pragma unroll
                        for (int i = 3; i < 6; ++i) lds128(Bf[2 * i], Bf[2 * i + 1], frag + i * 512);
                    }
                    if (s == 4) {
#pragma unroll
                        for (int i = 6; i < 8; ++i) lds128(Bf[2 * i], Bf[2 * i + 1], frag + i * 512);
                    }
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int t = 0; t < 3; ++t) dmma(acc[j][t][0], acc[j][t][1], cs.A[j][s], Bf[s * 3 + t]);
                }
                bar_arrive64(bar_pass_id);
                SR_TL(4);
            };
            auto probe_next = [&]() {   // the TRYWAIT latency hides behind this step's epilogue
                const bool wrap = (ring == NSTAGE - 1);
                next_ready = mbar_test_u32(wrap ? full_addr - 8 * (NSTAGE - 1) : full_addr + 8, wrap ? phase ^ 1 : phase);
            };
            // Any step, decoded at run time: the rare shapes (polytomies, a tip as root child, a gather that extends a partial
            // before its contraction). Written for few live registers, not speed: one operand at a time, stage held to the end.
            auto general_body = [&]() {
                const int n_pre = (rec.x >> CR_NPRE_SHIFT) & 3, mv = (rec.x >> CR_MV_SHIFT) & 3;
                const bool set0 = (rec.x & CR_PRE_SET) != 0, post = (rec.x & CR_POST) != 0;
                double t[M][5];
                uint32_t ix[M];
                for (int g = 0; g < n_pre; ++g) {
                    const uint32_t off = g == 0 ? rec.y : rec.z;
                    if (g == 0) { gather_idx(ix, 0); gather_rows(t, ix, 0, off); } else { gather_idx(ix, 1); gather_rows(t, ix, 1, off); }
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) cs.A[j][r] = (g == 0 && set0) ? t[j][r] : cs.A[j][r] * t[j][r];
                }
                if (mv) {
                    double Bf[16], acc[M][3][2];
                    const uint32_t frag = stg + frag_lane;
#pragma unroll
                    for (int i = 0; i < 3; ++i) lds128(Bf[2 * i], Bf[2 * i + 1], frag + i * 512);
                    burst(acc, Bf, frag);
                    probe_next();
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) cs.A[j][r] = acc[j][r >> 1][r & 1];
                    if (mv == 2) {
                        pop(t);
#pragma unroll
                        for (int j = 0; j < M; ++j)
#pragma unroll
                            for (int r = 0; r < 5; ++r) cs.A[j][r] *= t[j][r];
                    }
                    if (post) {
                        const uint32_t off = n_pre == 0 ? rec.y : (n_pre == 1 ? rec.z : rec.w);
                        if (n_pre == 0) { gather_idx(ix, 0); gather_rows(t, ix, 0, off); }
                        else if (n_pre == 1) { gather_idx(ix, 1); gather_rows(t, ix, 1, off); }
                        else { gather_idx(ix, 2); gather_rows(t, ix, 2, off); }
#pragma unroll
                        for (int j = 0; j < M; ++j)
#pragma unroll
                            for (int r = 0; r < 5; ++r) cs.A[j][r] *= t[j][r];
                    }
                }
                release(dep_of_A());          // A depends on every accumulator chain and on every gather
            };
            // One step of a shape known at compile time.
            auto body = [&](auto pre_c, auto mv_c, auto post_c) {
                constexpr int pre = decltype(pre_c)::value, mv = decltype(mv_c)::value;
                constexpr bool post = decltype(post_c)::value != 0;
                constexpr int n_pre = (pre == PRE_NEW2) ? 2 : (pre == PRE_NONE ? 0 : 1);
                double Bf[16], pv[M][5], vp[M][5], tp0[M][5], tp1[M][5];
                // ---- every load of the step is issued up front, longest latency first: the gathers (a clade's row comes from
                //      L1/L2), then the fragments of the first chunk, then the popped sibling
                uint32_t ix0[M], ix1[M], ix2[M];
                if (n_pre >= 1) gather_idx(ix0, 0);
                if (n_pre == 2) gather_idx(ix1, 1);
                if (post) gather_idx(ix2, n_pre);
                if (n_pre >= 1) gather_rows(tp0, ix0, 0, rec.y);
                if (n_pre == 2) gather_rows(tp1, ix1, 1, rec.z);
                if (post) gather_rows(vp, ix2, n_pre, n_pre == 0 ? rec.y : (n_pre == 1 ? rec.z : rec.w));
                const uint32_t frag = stg + frag_lane;
                if (mv) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) lds128(Bf[2 * i], Bf[2 * i + 1], frag + i * 512);
                }
                if (mv == 2) pop(pv);
                // ---- products that start or extend the partial
                if (pre == PRE_NEW2) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) cs.A[j][r] = tp0[j][r] * tp1[j][r];
                }
                double acc[M][3][2];
                burst(acc, Bf, frag);
                // Stage release. Without POST everything read from the stage has fed the accumulators: three chains of column tile
                // 0 between them consume every fragment, one chain of each further tile its gathers. A POST gather from a staged
                // tip table is only consumed by the product below, so with POST the release follows that product.
                if (!post) {
                    int dep = __double2hiint(acc[0][0][0]) | __double2hiint(acc[0][1][0]) | __double2hiint(acc[0][2][0]);
#pragma unroll
                    for (int j = 1; j < M; ++j) dep |= __double2hiint(acc[j][0][0]);
                    release(dep);
                }
                probe_next();
                SR_TL(5);
                if (mv == 2 && post) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) cs.A[j][r] = (acc[j][r >> 1][r & 1] * pv[j][r]) * vp[j][r];
                } else if (mv == 2) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) cs.A[j][r] = acc[j][r >> 1][r & 1] * pv[j][r];
                } else if (post) {
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) cs.A[j][r] = acc[j][r >> 1][r & 1] * vp[j][r];
                }
                if (post) release(dep_of_A());
                SR_TL(6);
                if (mv == 1 && !post) {
                    // a plain contraction has no product that moves its result into the operand registers for free
#pragma unroll
                    for (int j = 0; j < M; ++j)
#pragma unroll
                        for (int r = 0; r < 5; ++r) cs.A[j][r] = acc[j][r >> 1][r & 1];
                }
            };
#define SR_BODY(PRE, MV, POST) body(std::integral_constant<int, PRE>{}, std::integral_constant<int, MV>{}, std::integral_constant<int, POST>{})
            switch (rec.x & CR_SHAPE_MASK) {
                case SH_MV1: SR_BODY(PRE_NONE, 1, 0); break;
                case SH_MV2: SR_BODY(PRE_NONE, 2, 0); break;
                case SH_MV1_POST: SR_BODY(PRE_NONE, 1, 1); break;
                case SH_NEW2_MV1: SR_BODY(PRE_NEW2, 1, 0); break;
                case SH_NEW2_MV1_POST: SR_BODY(PRE_NEW2, 1, 1); break;
                case SH_NEW2_MV2: SR_BODY(PRE_NEW2, 2, 0); break;
                case SH_MV2_POST: SR_BODY(PRE_NONE, 2, 1); break;
                default: general_body();
            }
#undef SR_BODY
            if (rec.x & CR_TAIL) tail();
            SR_TL(7);
            // next ring position
            if (ring == NSTAGE - 1) { ring = 0; stage_addr -= (NSTAGE - 1) * P2_STAGE_BYTES; full_addr -= 8 * (NSTAGE - 1); phase ^= 1; }
            else { ++ring; stage_addr += P2_STAGE_BYTES; full_addr += 8; }
#ifdef SR_TIMELINE
            ++it;
#endif
        }
    }
}

// ------------------------------------------------------------------------------------------- K0b
// Precombined clade indices: for clade gather c of the schedule (tips' alignment rows ra, rb[, rc]) and every column of the
// launch, the row of its K1b table, (s_a*21 + s_b)[*21 + s_c], as uint16. One pass over the rows involved; prune_kernel
// then fetches one index per column and gather instead of two or three state bytes and the index arithmetic.
__global__ void __launch_bounds__(256) clade_index_kernel(const uint8_t* __restrict__ states, int64_t pitch, int64_t col0,
                                                          const int4* __restrict__ gathers, int64_t ncp,
                                                          uint16_t* __restrict__ cidx) {
    const int4 gt = gathers[blockIdx.x];                         // {row a, row b, row c or -1, -}
    const uint8_t* ra = states + col0 + (int64_t)gt.x * pitch;
    const uint8_t* rb = states + col0 + (int64_t)gt.y * pitch;
    const uint8_t* rc = gt.z >= 0 ? states + col0 + (int64_t)gt.z * pitch : nullptr;
    uint32_t* out = reinterpret_cast<uint32_t*>(cidx + (int64_t)blockIdx.x * ncp);
    // two columns per thread: 16-bit loads, one 32-bit store
    for (int64_t c2 = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; 2 * c2 < ncp; c2 += (int64_t)gridDim.y * blockDim.x) {
        const uint32_t a = *reinterpret_cast<const uint16_t*>(ra + 2 * c2), b = *reinterpret_cast<const uint16_t*>(rb + 2 * c2);
        uint32_t i0 = (a & 0xff) * 21 + (b & 0xff), i1 = (a >> 8) * 21 + (b >> 8);
        if (rc) {
            const uint32_t c = *reinterpret_cast<const uint16_t*>(rc + 2 * c2);
            i0 = i0 * 21 + (c & 0xff);
            i1 = i1 * 21 + (c >> 8);
        }
        out[c2] = i0 | (i1 << 16);
    }
}

}  // namespace srk
