// Does the FP64 tensor pipe keep its peak with the pruning kernel's real operand pattern (30 DMMAs per step over 10 distinct
// A registers, 15 distinct B registers, 6 accumulator pairs, zero-initialised every step) and with the Hadamard DMULs
// of the epilogue in between?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_pattern dmma_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NMUL>
__global__ void k(double* out, const double* in, int iters) {
    double A[2][5], Bf[15];
    for (int j = 0; j < 2; ++j) for (int r = 0; r < 5; ++r) A[j][r] = in[threadIdx.x + 32 * (j * 5 + r)];
    for (int i = 0; i < 15; ++i) Bf[i] = in[512 + threadIdx.x + 32 * i];
    double pv[2][5];
    for (int j = 0; j < 2; ++j) for (int r = 0; r < 5; ++r) pv[j][r] = 1.0 + 1e-9 * (threadIdx.x + r);
    for (int it = 0; it < iters; ++it) {
        double acc[2][3][2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int t = 0; t < 3; ++t) { acc[j][t][0] = 0.0; acc[j][t][1] = 0.0; }
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int t = 0; t < 3; ++t) dmma(acc[j][t][0], acc[j][t][1], A[j][s], Bf[s * 3 + t]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                double o = acc[j][r >> 1][r & 1];
                if (NMUL >= 1) o *= pv[j][r];
                if (NMUL >= 2) o *= pv[j][(r + 1) % 5];
                if (NMUL >= 3) o *= pv[1 - j][r];
                A[j][r] = o * 0.05 + 0.01;      // keep values bounded (1 more FP64 op: DFMA)
            }
    }
    double s = 0;
    for (int j = 0; j < 2; ++j) for (int r = 0; r < 5; ++r) s += A[j][r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NMUL>
void run(int threads, double* out, double* in) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<NMUL><<<148, threads>>>(out, in, iters);
    cudaEventRecord(e0); k<NMUL><<<148, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 256 * (148.0 * threads / 32) * iters * 30;
    printf("epilogue FP64 ops/elem=%d warps/SMSP=%.2f: %.2f TFLOP/s issued (x20/24 = %.2f useful)\n", NMUL + 1, threads / 128.0, fl / ms * 1e-9, fl / ms * 1e-9 * 20 / 24);
}
int main() {
    double *out, *in; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&in, 8 * 2048);
    cudaMemset(in, 0, 8 * 2048);
    for (int t : {128, 256, 384, 512}) run<0>(t, out, in);
    for (int t : {128, 256, 384, 512}) run<1>(t, out, in);
    for (int t : {256, 384}) run<3>(t, out, in);
    return 0;
}
