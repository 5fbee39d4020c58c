// FP64 roofline microbenchmarks for B200 (sm_100a): DFMA chain peak, DMMA (mma.sync f64) peak in
// every legal shape, and a DFMA + broadcast-LDS.128 mix. MEASURED_PEAKS.json has no FP64 entry, so
// bench.py's roofline denominator for the pruning kernel comes from this program's output
// (profiles/fp64_peak_*.json). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s @%d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 123.456) out[0] = s;
}

// mma.sync m8n8k4 f64: 256 FMA per warp instruction
template <int ACCS>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c0[ACCS], c1[ACCS];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
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

// m16n8k4: A 2 regs, B 1 reg, C 4 regs: 512 FMA
template <int ACCS>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b) {
    double c[ACCS][4];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ACCS; ++i)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C 4 regs: 1024 FMA
template <int ACCS>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[ACCS][4];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ACCS; ++i)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs: 2048 FMA
template <int ACCS>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double c[ACCS][4];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < ACCS; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// DFMA fed by broadcast LDS.128: per LDS.128 (2 doubles, same address for all lanes) do 2*COLS DFMA.
template <int COLS>
__global__ void __launch_bounds__(256) k_dfma_lds(double* out, int iters) {
    __shared__ double2 p[200];   // one 20x20 matrix
    for (int i = threadIdx.x; i < 200; i += blockDim.x) p[i] = make_double2(1.0 + 1e-9 * i, 1.0 - 1e-9 * i);
    __syncthreads();
    double acc[COLS][4];
    double x[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) { x[c] = 1.0 + threadIdx.x * 1e-6 + c; for (int r = 0; r < 4; ++r) acc[c][r] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 5
        for (int j = 0; j < 200; j += 2) {
            double2 a = p[j], b = p[j + 1];
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                acc[c][0] = fma(a.x, x[c], acc[c][0]);
                acc[c][1] = fma(a.y, x[c], acc[c][1]);
                acc[c][2] = fma(b.x, x[c], acc[c][2]);
                acc[c][3] = fma(b.y, x[c], acc[c][3]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < COLS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int sms = pr.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    const int T = 256;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", pr.name, sms, pr.clockRate);
    // ---- DFMA
    for (int bps : {2, 4, 8}) {
        int blocks = sms * bps, iters = 4096;
        {   double ms = time_ms([&] { k_dfma<8><<<blocks, T>>>(out, iters, 0.999999, 1e-7); }, 5);
            double fl = 2.0 * blocks * T * (double)iters * 8 * 8;
            printf(",\n \"dfma_chains8_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
        {   double ms = time_ms([&] { k_dfma<16><<<blocks, T>>>(out, iters, 0.999999, 1e-7); }, 5);
            double fl = 2.0 * blocks * T * (double)iters * 8 * 16;
            printf(",\n \"dfma_chains16_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
    }
    // ---- sustained DFMA (about 2 s) for the power-capped figure
    {
        int blocks = sms * 8, iters = 4096;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0; for (; n < 60; ++n) k_dfma<16><<<blocks, T>>>(out, iters, 0.999999, 1e-7);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * blocks * T * (double)iters * 8 * 16 * n;
        printf(",\n \"dfma_sustained_tflops\": %.3f, \"dfma_sustained_ms\": %.1f", fl / ms * 1e-9, ms);
    }
    // ---- DMMA shapes
    for (int bps : {1, 2, 4}) {
        int blocks = sms * bps, iters = 2048;
        int warps = blocks * T / 32;
        {   double ms = time_ms([&] { k_dmma884<8><<<blocks, T>>>(out, iters, 0.5, 0.25); }, 5);
            double fl = 2.0 * 256 * (double)warps * iters * 4 * 8;
            printf(",\n \"dmma_m8n8k4_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
        {   double ms = time_ms([&] { k_dmma1684<4><<<blocks, T>>>(out, iters, 0.5, 0.25); }, 5);
            double fl = 2.0 * 512 * (double)warps * iters * 4 * 4;
            printf(",\n \"dmma_m16n8k4_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
        {   double ms = time_ms([&] { k_dmma1688<4><<<blocks, T>>>(out, iters, 0.5, 0.25); }, 5);
            double fl = 2.0 * 1024 * (double)warps * iters * 4 * 4;
            printf(",\n \"dmma_m16n8k8_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
        {   double ms = time_ms([&] { k_dmma16816<4><<<blocks, T>>>(out, iters, 0.5, 0.25); }, 5);
            double fl = 2.0 * 2048 * (double)warps * iters * 2 * 4;
            printf(",\n \"dmma_m16n8k16_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
    }
    // ---- DFMA + broadcast LDS
    for (int bps : {1, 2, 4}) {
        int blocks = sms * bps, iters = 256;
        {   double ms = time_ms([&] { k_dfma_lds<1><<<blocks, T>>>(out, iters); }, 5);
            double fl = 2.0 * blocks * T * (double)iters * 400 * 1;
            printf(",\n \"dfma_lds128_cols1_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
        {   double ms = time_ms([&] { k_dfma_lds<2><<<blocks, T>>>(out, iters); }, 5);
            double fl = 2.0 * blocks * T * (double)iters * 400 * 2;
            printf(",\n \"dfma_lds128_cols2_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
        {   double ms = time_ms([&] { k_dfma_lds<4><<<blocks, T>>>(out, iters); }, 5);
            double fl = 2.0 * blocks * T * (double)iters * 400 * 4;
            printf(",\n \"dfma_lds128_cols4_bps%d_tflops\": %.3f", bps, fl / ms * 1e-9); }
    }
    printf("\n}\n");
    return 0;
}
