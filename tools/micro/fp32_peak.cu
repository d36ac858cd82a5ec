// fp32_peak.cu — microbenchmarks behind the FP32 roofline of the f32 pair kernel: sustained rates of scalar FFMA /
// FADD, of the packed FFMA2 / FADD2 forms of sm_100 (fma.rn.f32x2 / add.rn.f32x2), of MUFU.EX2, and of the mixes
// the kernel issues (d subtractions + d FMAs + 1 ex2 + 1 add per pair).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp32_peak tools/micro/fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), #x); return 1; } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { return (u64)__float_as_uint(a) | ((u64)__float_as_uint(b) << 32); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// MODE 0 FFMA, 1 FADD, 2 FFMA2, 3 FADD2, 4 FFMA+FADD 1:1, 5 FFMA2+FADD2 1:1, 6 MUFU.EX2, 7 FFMA:FADD:EX2 = 4:5:1 (scalar
// pair-kernel mix, d = 4), 8 FFMA2:FADD2:EX2 = 4:5:2 (packed mix: 2 pairs per step)
template <int CHAINS, int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[CHAINS], y[CHAINS];
    u64 p[CHAINS], q[CHAINS];
    const u64 pa = pack(a, a), pb = pack(b, b);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = threadIdx.x * 1e-6f + i; y[i] = x[i] + 1.f; p[i] = pack(x[i], y[i]); q[i] = pack(y[i], x[i]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);
            if (MODE == 1) x[i] = x[i] + b;
            if (MODE == 2) p[i] = ffma2(p[i], pa, pb);
            if (MODE == 3) p[i] = fadd2(p[i], pb);
            if (MODE == 4) { x[i] = fmaf(x[i], a, b); y[i] = y[i] + b; }
            if (MODE == 5) { p[i] = ffma2(p[i], pa, pb); q[i] = fadd2(q[i], pb); }
            if (MODE == 6) x[i] = ex2(x[i]);
            if (MODE == 7) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) { float d = x[i] - (b + c); acc = fmaf(-d, d, acc); }
                y[i] += ex2(acc);
            }
            if (MODE == 8) {
                u64 acc = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) { u64 d = fadd2(p[i], pack(-(b + c), -(b + c))); acc = ffma2(d ^ 0x8000000080000000ull, d, acc); }
                float lo = ex2(__uint_as_float((unsigned)acc)), hi = ex2(__uint_as_float((unsigned)(acc >> 32)));
                q[i] = fadd2(q[i], pack(lo, hi));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i] + y[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)q[i]) + __uint_as_float((unsigned)(q[i] >> 32));
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
int run(const char* name, double ops_per_chain_step, int sms, int warps_per_sm) {
    constexpr int CH = 8;
    float* d; CHECK(cudaMalloc(&d, 4));
    int iters = 1 << 15;
    int blocks = sms * warps_per_sm / 8;
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    k<CH, MODE><<<blocks, 256>>>(d, 64, 1.0001f, 0.5f);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    k<CH, MODE><<<blocks, 256>>>(d, iters, 1.0001f, 0.5f);
    CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    double steps = (double)blocks * 256 * CH * iters;
    double per_clk_sm = steps * ops_per_chain_step / (ms * 1e-3) / 1.965e9 / sms;
    printf("warps/SM %2d  %-34s: %8.3f ms  %7.2f lane-ops/clk/SM  (%.2f chain-steps/clk/SM)\n", warps_per_sm, name, ms, per_clk_sm,
           steps / (ms * 1e-3) / 1.965e9 / sms);
    cudaFree(d);
    return 0;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs\n", p.name, sms);
    for (int w : {16, 32}) {
        run<0>("FFMA scalar", 1, sms, w);
        run<1>("FADD scalar", 1, sms, w);
        run<2>("FFMA2 packed (2 FMA/lane)", 2, sms, w);
        run<3>("FADD2 packed (2 add/lane)", 2, sms, w);
        run<4>("FFMA + FADD 1:1 scalar", 2, sms, w);
        run<5>("FFMA2 + FADD2 1:1 packed", 4, sms, w);
        run<6>("MUFU.EX2", 1, sms, w);
        run<7>("pair mix d=4 scalar (1 pair/step)", 1, sms, w);
        run<8>("pair mix d=4 packed (2 pairs/step)", 2, sms, w);
    }
    return 0;
}
