// Does a DMMA burst of one warp leave the sub-partition's issue port to the other warps?  Warps 0..3 (one per
// sub-partition) issue nothing but DMMAs (6 independent accumulator chains, the pruning kernel's pattern); warps 4..7
// issue a stream of other instructions (integer, FP64 multiply, or shared-memory loads) until the DMMA warps are done.
// Reported: cycles per DMMA seen by the MMA warp, and instructions per cycle achieved by the other warp next to it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_port dmma_port.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: no companion work, 1: integer IMAD, 2: DMUL, 3: LDS.64, 4: IMAD but only `duty` of every 16 slots
template <int MODE>
__global__ void k(double* out, long long* stats, int iters, int duty) {
    __shared__ double sm[1024];
    __shared__ volatile int done;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 4) {
        double acc[6][2] = {};
        double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
                for (int c = 0; c < 6; ++c) dmma(acc[c][0], acc[c][1], a, b);
        }
        long long t1 = clock64();
        double s = 0;
        for (int c = 0; c < 6; ++c) s += acc[c][0] + acc[c][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
        if (lane == 0) { stats[(blockIdx.x * 8 + warp) * 2] = t1 - t0; atomicAdd((int*)&done, 1); }
    } else {
        unsigned x[8];
        double d[8];
        for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 7 + i; d[i] = 1.0 + i * 1e-3; }
        long long n = 0;
        long long t0 = clock64();
        if (MODE != 0) {
            while (done < 4) {
#pragma unroll
                for (int rep = 0; rep < 4; ++rep) {
                    if (MODE == 1 || MODE == 4) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(x[i]));
                    } else if (MODE == 2) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) asm volatile("mul.f64 %0, %0, 1.0000001;" : "+d"(d[i]));
                    } else if (MODE == 3) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(lane + 32 * i) & 1023]);
                            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(d[i]) : "r"(addr));
                        }
                    }
                    if (MODE == 4) __nanosleep(duty);
                }
                n += 32;
            }
        }
        long long t1 = clock64();
        double s = 0;
        for (int i = 0; i < 8; ++i) s += d[i] + x[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
        if (lane == 0) { stats[(blockIdx.x * 8 + warp) * 2] = t1 - t0; stats[(blockIdx.x * 8 + warp) * 2 + 1] = n; }
    }
}
template <int MODE>
void run(const char* name, double* out, long long* stats, int duty = 0) {
    const int iters = 20000;
    long long h[148 * 16];
    k<MODE><<<148, 256>>>(out, stats, iters, duty);
    cudaMemset(stats, 0, sizeof(h));
    k<MODE><<<148, 256>>>(out, stats, iters, duty);
    cudaMemcpy(h, stats, sizeof(h), cudaMemcpyDeviceToHost);
    double mma = 0, oth = 0, othc = 0;
    for (int b = 0; b < 148; ++b) for (int w = 0; w < 4; ++w) {
        mma += (double)h[(b * 8 + w) * 2];
        oth += (double)h[(b * 8 + w + 4) * 2 + 1];
        othc += (double)h[(b * 8 + w + 4) * 2];
    }
    printf("%-28s cycles per DMMA %.2f   companion instructions per cycle %.3f\n", name, mma / (148.0 * 4 * iters * 30), othc > 0 ? oth / othc : 0.0);
}
int main() {
    double* out; long long* stats;
    cudaMalloc(&out, 8 * 148 * 256); cudaMalloc(&stats, 8 * 148 * 16);
    run<0>("MMA warp alone", out, stats);
    run<1>("+ IMAD warp flat out", out, stats);
    run<2>("+ DMUL warp flat out", out, stats);
    run<3>("+ LDS.64 warp flat out", out, stats);
    run<4>("+ IMAD warp, nanosleep 20", out, stats, 20);
    run<4>("+ IMAD warp, nanosleep 100", out, stats, 100);
    run<4>("+ IMAD warp, nanosleep 400", out, stats, 400);
    return 0;
}
