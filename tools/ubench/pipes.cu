// pipes.cu -- which SM pipes overlap with packed FP32 (FFMA2) on sm_100a?  Development micro-benchmark:
// every kernel runs NF FFMA2 + NO "other" instructions per inner step, 8 independent chains each, 16 warps per
// SMSP; reports SMSP cycles per inner step assuming FFMA2 alone costs 2 cycles per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long pk64;
__device__ __forceinline__ pk64 pk(float x, float y) { pk64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }

enum { O_NONE, O_LOP3, O_IADD3, O_IMAD, O_IMADW, O_LDS, O_MUFU, O_FFMA, O_SHF, O_ISETP, O_STS, O_PRMT, O_I2F, O_F2I, O_FADD2, O_LDS128, O_FMNMX, O_SEL };

template <int OTHER, int NF, int NO>
__global__ void __launch_bounds__(256) mix(float* out, int iters, float a, float b, unsigned k) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = i * 1e-3f;
    __syncthreads();
    pk64 v[8], ab = pk(a, a), bb = pk(b, b);
    unsigned z[8]; float f[8]; unsigned long long w[8];
    float4 q4 = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = pk((threadIdx.x + i) * 1e-3f, 1.0f); z[i] = threadIdx.x * 2654435761u + i; f[i] = 1.0f + i * 1e-3f + threadIdx.x * 1e-4f; w[i] = z[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < NF) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(ab), "l"(bb));
                if (i < NO) {
                    if (OTHER == O_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(k), "r"(z[(i + 1) & 7]));
                    if (OTHER == O_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(z[(i + 1) & 7]));
                    if (OTHER == O_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(k), "r"(z[(i + 1) & 7]));
                    if (OTHER == O_IMADW) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((unsigned)w[i]), "r"(k));
                    if (OTHER == O_LDS) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[i]) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((z[i] + r * 4) & 8188u)));
                    if (OTHER == O_LDS128) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q4.x), "=f"(q4.y), "=f"(q4.z), "=f"(f[i]) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((threadIdx.x * 16 + r * 512 + i * 64) & 8176u)));
                    if (OTHER == O_STS) asm volatile("st.shared.f32 [%1], %0;" :: "f"(f[i]), "r"((unsigned)__cvta_generic_to_shared(sm) + ((threadIdx.x * 4 + r * 1024 + i * 128) & 8188u)));
                    if (OTHER == O_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                    if (OTHER == O_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
                    if (OTHER == O_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(z[i]) : "r"(z[(i + 1) & 7]));
                    if (OTHER == O_ISETP) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, 1;}" : "+r"(z[i]) : "r"(z[(i + 1) & 7]));
                    if (OTHER == O_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(z[i]) : "r"(z[(i + 1) & 7]));
                    if (OTHER == O_I2F) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(z[i]));
                    if (OTHER == O_F2I) asm volatile("cvt.rpi.u32.f32 %0, %1;" : "=r"(z[i]) : "f"(f[i]));
                    if (OTHER == O_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(bb));
                    if (OTHER == O_FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
                    if (OTHER == O_SEL) asm volatile("slct.u32.s32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(z[(i + 1) & 7]), "r"((int)z[(i + 2) & 7]));
                }
            }
        }
    }
    float s = q4.x + q4.y + q4.z; unsigned zz = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i])); s += lo + hi + f[i]; zz ^= z[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32); }
    if (s == 12345.678f || zz == 0x12345u) out[0] = s;
}

template <int OTHER, int NF, int NO> float run(float* out, int grid, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix<OTHER, NF, NO><<<grid, 256>>>(out, iters, 0.999f, 1e-4f, 0x9E3779B9u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}
#define ROW(name, O)                                                                                   \
    { float t8 = run<O, 8, 8>(out, grid, iters), t4 = run<O, 8, 4>(out, grid, iters), t0 = run<O, 0, 8>(out, grid, iters), t2 = run<O, 8, 2>(out, grid, iters); \
      printf("%-8s alone %.2f cyc/inst | with 8 FFMA2: +8 other -> %.2f cyc per (FFMA2+other), +4 -> %.2f per FFMA2, +2 -> %.2f per FFMA2\n", name, \
             t0 / base * 2.0, t8 / base * 2.0, t4 / base * 2.0, t2 / base * 2.0); }
int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    float* out; cudaMalloc(&out, 64);
    const int grid = p.multiProcessorCount * 8, iters = 2048;
    float base = run<O_NONE, 8, 0>(out, grid, iters);   // 64 FFMA2 per iteration = 128 SMSP cycles at 2 cyc each... per warp, 16 warps/SMSP
    printf("FFMA2 only: %.3f ms (%.1f TFLOP/s)\n", base, 2.0 * 128 * iters * 256.0 * grid / (base * 1e-3) / 1e12);
    ROW("LOP3", O_LOP3) ROW("IADD3", O_IADD3) ROW("IMAD", O_IMAD) ROW("IMAD.W", O_IMADW) ROW("LDS", O_LDS) ROW("LDS128", O_LDS128) ROW("STS", O_STS)
    ROW("MUFU", O_MUFU) ROW("FFMA", O_FFMA) ROW("SHF", O_SHF) ROW("ISETP+@", O_ISETP) ROW("PRMT", O_PRMT) ROW("I2F", O_I2F) ROW("F2I", O_F2I)
    ROW("FADD2", O_FADD2) ROW("FMNMX", O_FMNMX) ROW("SLCT", O_SEL)
    return 0;
}
