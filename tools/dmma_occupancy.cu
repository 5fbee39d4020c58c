// How many warps per SM sub-partition does the FP64 tensor pipe need to saturate, at 6 accumulator chains per warp
// (what one pruning step offers with M=2)? Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_occupancy dmma_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ACCS>
__global__ void k(double* out, int iters, double a, double b) {
    double c0[ACCS], c1[ACCS];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int i = 0; i < ACCS; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}
template <int ACCS>
void run(int threads, double* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096;
    for (int i = 0; i < 2; ++i) k<ACCS><<<148, threads>>>(out, iters, 0.5, 0.25);
    cudaEventRecord(e0); k<ACCS><<<148, threads>>>(out, iters, 0.5, 0.25); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 256 * (148.0 * threads / 32) * iters * 5 * ACCS;
    printf("chains/warp=%d warps/SMSP=%.2f: %.2f TFLOP/s\n", ACCS, threads / 128.0, fl / ms * 1e-9);
}
int main() {
    double* out; cudaMalloc(&out, 8);
    for (int t : {128, 256, 384, 512}) run<6>(t, out);
    for (int t : {128, 256}) run<3>(t, out);
    for (int t : {128, 256}) run<12>(t, out);
    return 0;
}
