// Can a warp issue other work between the DMMAs of its burst without slowing the FP64 tensor pipe? K integer ops (or one
// LDS.64 per KL DMMAs) are interleaved between consecutive DMMAs of the pruning kernel's 30-DMMA step pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_coissue dmma_coissue.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int KINT, int LDS_EVERY>
__global__ void k(double* out, const double* in, int iters) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    double A[2][5], Bf[15];
    for (int j = 0; j < 2; ++j) for (int r = 0; r < 5; ++r) A[j][r] = in[threadIdx.x % 32 + 32 * (j * 5 + r)];
    for (int i = 0; i < 15; ++i) Bf[i] = in[512 + threadIdx.x % 32 + 32 * i];
    unsigned x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 7 + i;
    double ld = 0;
    for (int it = 0; it < iters; ++it) {
        double acc[2][3][2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int t = 0; t < 3; ++t) { acc[j][t][0] = 0.0; acc[j][t][1] = 0.0; }
        int n = 0;
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    dmma(acc[j][t][0], acc[j][t][1], A[j][s], Bf[s * 3 + t]);
#pragma unroll
                    for (int q = 0; q < KINT; ++q) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(x[q % 8]));
                    if (LDS_EVERY > 0 && (n % LDS_EVERY) == 0) {
                        double v; unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x * 3 + n * 32) & 2047]);
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
                        ld += v;
                    }
                    ++n;
                }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 5; ++r) A[j][r] = acc[j][r >> 1][r & 1] * 0.05 + 0.01;
    }
    double s = ld;
    for (int j = 0; j < 2; ++j) for (int r = 0; r < 5; ++r) s += A[j][r];
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KINT, int LDS_EVERY>
void run(int threads, double* out, double* in) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<KINT, LDS_EVERY><<<148, threads>>>(out, in, iters);
    cudaEventRecord(e0); k<KINT, LDS_EVERY><<<148, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 256 * (148.0 * threads / 32) * iters * 30;
    printf("int ops per DMMA=%2d, LDS every %d DMMAs, warps/SMSP=%.0f: %.2f TFLOP/s issued\n", KINT, LDS_EVERY, threads / 128.0, fl / ms * 1e-9);
}
int main() {
    double *out, *in; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&in, 8 * 2048); cudaMemset(in, 0, 8 * 2048);
    run<0, 0>(128, out, in); run<2, 0>(128, out, in); run<4, 0>(128, out, in); run<8, 0>(128, out, in); run<12, 0>(128, out, in); run<16, 0>(128, out, in);
    run<0, 1>(128, out, in); run<4, 1>(128, out, in); run<0, 2>(128, out, in);
    run<8, 0>(256, out, in); run<8, 1>(256, out, in); run<8, 1>(384, out, in);
    return 0;
}
