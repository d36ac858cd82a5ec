// fp64_mix.cu - what one non-FP64 instruction costs when it is interleaved with a DFMA stream on sm_100a (round 2).
// The f64 pair kernel issues 17 FP64 + ~12 other instructions per row pair; profiles/r1c_tuning.md showed that the classes
// differ (IMAD ~1 clk, LOP3 ~0.3 clk per DFMA pair).  This benchmark puts a number on every class the exp2 glue can be built
// from, at the kernel's own ratio (NI others per 6 DFMA), with the loop unrolled so that loop overhead is < 2%.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_mix tools/micro/fp64_mix.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), #x); return 1; } } while (0)

enum { NONE = 0, LOP3, SHF, IADD3, IMAD, IMADSHL, VIADDMNMX, VIMNMX, LEA, PRMT, LDS_BCAST, LDS_LANE, FFMA, MOV_IMM, NMODES };
static const char* kNames[NMODES] = {"none", "LOP3", "SHF", "add + LOP3", "IMAD", "shl + LOP3", "VIADDMNMX", "VIMNMX", "LEA", "PRMT",
                                     "LDS.64 (one address)", "LDS.64 (lane stride 8B)", "FFMA", "LOP3 imm"};

template <int MODE>
__device__ __forceinline__ void other(int& y, int m, int it, float& f, double& acc, const double* sh, uint32_t sh_base) {
    if (MODE == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(y) : "r"(m), "r"(it));
    if (MODE == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(y));
    if (MODE == IADD3) { int t; asm volatile("add.s32 %0, %1, %2;\n\txor.b32 %1, %0, %3;" : "=r"(t), "+r"(y) : "r"(m), "r"(it)); }  // + LOP3
    if (MODE == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y) : "r"(m), "r"(it));
    if (MODE == IMADSHL) { int t; asm volatile("shl.b32 %0, %1, 3;\n\tand.b32 %1, %0, %2;" : "=r"(t), "+r"(y) : "r"(m)); }  // + LOP3
    if (MODE == VIADDMNMX) { int t; asm volatile("add.s32 %0, %1, %2;\n\tmax.s32 %1, %0, %3;" : "=r"(t), "+r"(y) : "r"(m), "r"(it)); }
    if (MODE == VIMNMX) asm volatile("max.s32 %0, %0, %1;" : "+r"(y) : "r"(it));
    if (MODE == LEA) { int t; asm volatile("shl.b32 %0, %1, 8;\n\tadd.s32 %1, %0, %2;" : "=r"(t), "+r"(y) : "r"(m)); }
    if (MODE == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(y) : "r"(m));
    // the loaded value has no consumer: what is measured is the issue cost of the load next to the DFMA stream
    if (MODE == LDS_BCAST) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(sh_base + ((unsigned)(it & 0xff) << 3))); }
    if (MODE == LDS_LANE) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(sh_base + ((threadIdx.x & 31) << 3) + ((unsigned)(it & 0xf) << 8))); }
    if (MODE == FFMA) f = fmaf(f, 1.0001f, 0.5f);
    if (MODE == MOV_IMM) { int t; asm volatile("mov.b32 %0, 0x3f262e42;\n\txor.b32 %1, %1, %0;" : "=r"(t), "+r"(y)); }
}

// per step: 6 DFMA (one exp2's worth) and NI instructions of class MODE, 8 independent steps per iteration, x2 unrolled
template <int MODE, int NI>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b, int m) {
    __shared__ double sh[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sh[i] = i;
    __syncthreads();
    const uint32_t sh_base = static_cast<uint32_t>(__cvta_generic_to_shared(sh));
    constexpr int C = 8;
    double x[C];
    int y[C];
    float f[C];
    double acc[C];
#pragma unroll
    for (int i = 0; i < C; ++i) { x[i] = threadIdx.x * 1e-9 + i; y[i] = threadIdx.x + i; f[i] = i; acc[i] = 0; }
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                x[i] = fma(x[i], a, b);
                if (k < NI) other<MODE>(y[i], m, it, f[i], acc[i], sh, sh_base);
            }
        }
    }
    double s = 0; int t = 0; float g = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) { s += x[i] + acc[i]; t ^= y[i]; g += f[i]; }
    if (s == 123.456 || t == 0x7fffffff || g == 1.2345f) out[0] = s + t + g;
}


// ---- register-operand traffic of the FP64 instruction itself: how many 64-bit REGISTER sources it reads ----
// OPS: 0 DADD x += imm (1 register source)        1 DFMA x = x * c[param] + imm?? not encodable -> x = x * x + imm (1 distinct)
//      2 DFMA x = x * param + r (2: x and a loop-invariant register)   3 DFMA x = x * y_i + imm (2 distinct, y_i differs per chain)
//      4 DFMA x = x * y_i + z_i (3 distinct registers, no reuse)      5 DFMA x = x * r1 + r2 (3, two of them loop-invariant)
//      6 DADD x = x + y_i (2 distinct)
template <int OPS>
__global__ void __launch_bounds__(256) k_ops(double* out, const double* in, int iters, double a) {
    constexpr int C = 8;
    double x[C], y[C], z[C];
#pragma unroll
    for (int i = 0; i < C; ++i) { x[i] = threadIdx.x * 1e-9 + i; y[i] = in[i] + threadIdx.x * 1e-12; z[i] = in[C + i] + threadIdx.x * 1e-12; }
    const double r1 = in[2 * C] + threadIdx.x * 1e-12, r2 = in[2 * C + 1] + threadIdx.x * 1e-12;  // vector registers, not uniform ones
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
                if (OPS == 0) x[i] = x[i] + 6755399441055744.0;
                if (OPS == 1) x[i] = fma(x[i], x[i], 1.0);
                if (OPS == 2) x[i] = fma(x[i], a, r1);
                if (OPS == 3) x[i] = fma(x[i], y[i], 1.0);
                if (OPS == 4) x[i] = fma(x[i], y[i], z[i]);
                if (OPS == 5) x[i] = fma(x[i], r1, r2);
                if (OPS == 6) x[i] = x[i] + y[i];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}
template <int OPS>
static void ops_row(double* out, const double* in, int ctas, int iters, double ghz, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_ops<OPS><<<ctas, 256>>>(out, in, iters, 1.0000001);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_ops<OPS><<<ctas, 256>>>(out, in, iters, 1.0000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-60s %5.2f clk per instruction and scheduler\n", name, ms * 1e-3 * ghz * 1e9 / (2.0 * 48.0 * iters));
}

static float time_ms(void (*launch)(double*, int, int), double* out, int ctas, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(out, ctas, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(out, ctas, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

template <int MODE, int NI>
static void launch(double* out, int ctas, int iters) { k_mix<MODE, NI><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9, 0xff); }

template <int MODE>
static void row(double* out, int ctas, int iters, double clk_ghz, float base_ms) {
    float m1 = time_ms(launch<MODE, 1>, out, ctas, iters), m3 = time_ms(launch<MODE, 3>, out, ctas, iters),
          m5 = time_ms(launch<MODE, 5>, out, ctas, iters);
    // clocks per scheduler and step (6 DFMA + NI others): 2 warps per scheduler at 8 warps/SM
    const double steps = 2.0 * 8.0 * iters;
    auto clk = [&](float ms) { return ms * 1e-3 * clk_ghz * 1e9 / steps; };
    printf("%-26s 6 DFMA + 1: %6.2f clk  + 3: %6.2f clk  + 5: %6.2f clk   per extra instruction: %5.2f / %5.2f / %5.2f clk\n", kNames[MODE],
           clk(m1), clk(m3), clk(m5), clk(m1) - clk(base_ms), (clk(m3) - clk(base_ms)) / 3, (clk(m5) - clk(base_ms)) / 5);
}

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out;
    CHECK(cudaMalloc(&out, 8));
    const int iters = 1 << 13;
    const double ghz = p.clockRate * 1e-6;
    printf("device %s, %d SMs, %.3f GHz; 8 warps / SM (2 per scheduler), 8 independent chains per thread\n", p.name, sms, ghz);
    int ctas = sms;
    float base = time_ms(launch<NONE, 0>, out, ctas, iters);
    printf("6 DFMA alone: %.2f clk per scheduler (2.00 per DFMA = nominal)\n", base * 1e-3 * ghz * 1e9 / (2.0 * 8.0 * iters));
    row<LOP3>(out, ctas, iters, ghz, base);
    row<SHF>(out, ctas, iters, ghz, base);
    row<IADD3>(out, ctas, iters, ghz, base);
    row<IMAD>(out, ctas, iters, ghz, base);
    row<IMADSHL>(out, ctas, iters, ghz, base);
    row<VIADDMNMX>(out, ctas, iters, ghz, base);
    row<VIMNMX>(out, ctas, iters, ghz, base);
    row<LEA>(out, ctas, iters, ghz, base);
    row<PRMT>(out, ctas, iters, ghz, base);
    row<LDS_BCAST>(out, ctas, iters, ghz, base);
    row<LDS_LANE>(out, ctas, iters, ghz, base);
    row<FFMA>(out, ctas, iters, ghz, base);
    row<MOV_IMM>(out, ctas, iters, ghz, base);
    double* in;
    CHECK(cudaMalloc(&in, 64 * 8));
    double hin[64];
    for (int i = 0; i < 64; ++i) hin[i] = 1.0 + 1e-9 * i;
    CHECK(cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice));
    ops_row<0>(out, in, ctas, iters, ghz, "DADD x + imm                      (1 register source)");
    ops_row<1>(out, in, ctas, iters, ghz, "DFMA x * x + imm                  (1 distinct)");
    ops_row<2>(out, in, ctas, iters, ghz, "DFMA x * const + r                (2, one loop-invariant)");
    ops_row<3>(out, in, ctas, iters, ghz, "DFMA x * y_i + imm                (2 distinct)");
    ops_row<4>(out, in, ctas, iters, ghz, "DFMA x * y_i + z_i                (3 distinct)");
    ops_row<5>(out, in, ctas, iters, ghz, "DFMA x * r1 + r2                  (3, two loop-invariant)");
    ops_row<6>(out, in, ctas, iters, ghz, "DADD x + y_i                      (2 distinct)");
    return 0;
}
