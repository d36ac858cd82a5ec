// fp64_peak.cu — microbenchmarks behind the FP64 roofline of the pair kernel (SURVEY.md §8d asks
// for "a pure-FMA microbenchmark"): sustained DFMA rate, DMMA (m8n8k4) rate, and whether the two
// overlap when issued from the same SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o tools/micro/fp64_peak tools/micro/fp64_peak.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), #x); return 1; } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

// DFMA with an equal number of integer (IMAD/LOP) instructions interleaved
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma_int(double* out, int iters, double a, double b, int m) {
    double x[CHAINS];
    int y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = threadIdx.x * 1e-9 + i; y[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) { x[i] = fma(x[i], a, b); y[i] = y[i] * m + it; }
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { s += x[i]; t ^= y[i]; }
    if (s == 123.456 || t == 0x7fffffff) out[0] = s + t;
}

// DFMA interleaved 1:1 with another instruction class (MODE: 0 LOP3, 1 shift-add (LEA), 2 FFMA, 3 LDS, 4 IMAD every 4th)
template <int CHAINS, int MODE>
__global__ void __launch_bounds__(256) k_dfma_mix(double* out, int iters, double a, double b, int m, float fa) {
    __shared__ double sh[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = i;
    __syncthreads();
    double x[CHAINS];
    int y[CHAINS];
    float f[CHAINS];
    double acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = threadIdx.x * 1e-9 + i; y[i] = threadIdx.x + i; f[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            x[i] = fma(x[i], a, b);
            if (MODE == 0) y[i] = (y[i] & m) ^ it;
            if (MODE == 1) y[i] = (y[i] << 3) + it;
            if (MODE == 2) f[i] = fmaf(f[i], fa, fa);
            if (MODE == 3) { acc += sh[(y[i] + it) & 2047 & ~15 | (threadIdx.x & 15)]; }
            if (MODE == 4 && (i & 3) == 0) y[i] = y[i] * m + it;
        }
    }
    double s = acc; int t = 0; float g = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { s += x[i]; t ^= y[i]; g += f[i]; }
    if (s == 123.456 || t == 0x7fffffff || g == 1.2345f) out[0] = s + t + g;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ACCS>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
    double c[ACCS][2];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACCS; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// per iteration: ACCS DMMA + NF DFMA per accumulator slot
template <int ACCS, int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double c[ACCS][2];
    double x[ACCS * NF];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < ACCS * NF; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACCS; ++i) {
            dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < NF; ++j) x[i * NF + j] = fma(x[i * NF + j], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < ACCS * NF; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F&& launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out;
    CHECK(cudaMalloc(&out, 8));
    const int iters = 1 << 15;
    printf("device %s, %d SMs\n", p.name, sms);
    for (int wpsm : {8, 16}) {   // warps per SM
        int ctas = sms * wpsm / 8;
        double warps = (double)ctas * 8;
        {
            float ms = time_ms([&] { k_dfma<8><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); });
            double inst = warps * 32 * 8.0 * iters;
            printf("warps/SM %2d  DFMA x8 chains      : %8.3f ms  %6.2f TFLOP/s  %6.2f lane-FMA/clk/SM @1965MHz\n", wpsm, ms,
                   2 * inst / ms * 1e-9, inst / (ms * 1e-3) / sms / 1.965e9);
        }
        {
            float ms = time_ms([&] { k_dfma_int<8><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9, 3); });
            double inst = warps * 32 * 8.0 * iters;
            printf("warps/SM %2d  DFMA + IMAD (1:1)   : %8.3f ms  %6.2f TFLOP/s  %6.2f lane-FMA/clk/SM\n", wpsm, ms, 2 * inst / ms * 1e-9,
                   inst / (ms * 1e-3) / sms / 1.965e9);
        }
        {
            const char* names[5] = {"DFMA + LOP3 (1:1)", "DFMA + LEA  (1:1)", "DFMA + FFMA (1:1)", "DFMA + LDS  (1:1)", "DFMA + IMAD (4:1)"};
            float msv[5];
            msv[0] = time_ms([&] { k_dfma_mix<8, 0><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9, 0xff, 1.0001f); });
            msv[1] = time_ms([&] { k_dfma_mix<8, 1><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9, 0xff, 1.0001f); });
            msv[2] = time_ms([&] { k_dfma_mix<8, 2><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9, 0xff, 1.0001f); });
            msv[3] = time_ms([&] { k_dfma_mix<8, 3><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9, 0xff, 1.0001f); });
            msv[4] = time_ms([&] { k_dfma_mix<8, 4><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9, 3, 1.0001f); });
            double inst = warps * 32 * 8.0 * iters;
            for (int q = 0; q < 5; ++q)
                printf("warps/SM %2d  %s   : %8.3f ms  %6.2f lane-FMA/clk/SM\n", wpsm, names[q], msv[q], inst / (msv[q] * 1e-3) / sms / 1.965e9);
        }
        {
            float ms = time_ms([&] { k_dmma<8><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); });
            double fma_ = warps * 8.0 * iters * 256.0;  // 8x8x4 FMAs per warp-level DMMA
            printf("warps/SM %2d  DMMA m8n8k4 x8 accs : %8.3f ms  %6.2f TFLOP/s  %6.2f FMA/clk/SM\n", wpsm, ms, 2 * fma_ / ms * 1e-9,
                   fma_ / (ms * 1e-3) / sms / 1.965e9);
        }
        {
            float ms = time_ms([&] { k_mixed<4, 4><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); });
            double dm = warps * 4.0 * iters * 256.0, df = warps * 32 * 16.0 * iters;
            printf("warps/SM %2d  mixed 1 DMMA:4 DFMA : %8.3f ms  DMMA %6.2f + DFMA %6.2f TFLOP/s (DFMA %6.2f lane-FMA/clk/SM)\n", wpsm, ms,
                   2 * dm / ms * 1e-9, 2 * df / ms * 1e-9, df / (ms * 1e-3) / sms / 1.965e9);
        }
        {
            float ms = time_ms([&] { k_mixed<2, 8><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); });
            double dm = warps * 2.0 * iters * 256.0, df = warps * 32 * 16.0 * iters;
            printf("warps/SM %2d  mixed 1 DMMA:8 DFMA : %8.3f ms  DMMA %6.2f + DFMA %6.2f TFLOP/s (DFMA %6.2f lane-FMA/clk/SM)\n", wpsm, ms,
                   2 * dm / ms * 1e-9, 2 * df / ms * 1e-9, df / (ms * 1e-3) / sms / 1.965e9);
        }
    }
    return 0;
}
