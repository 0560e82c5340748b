// pipes2.cu -- second round of sm_100a pipe micro-benchmarks (development aid): operand-form sensitivity of
// FFMA2 co-issue, and the cost of one Philox4x32-10 call / one Box-Muller sample, alone and mixed with FFMA2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long pk64;
__device__ __forceinline__ pk64 pk(float x, float y);
__device__ __forceinline__ pk64 pk(float2 a) { return pk(a.x, a.y); }
__device__ __forceinline__ pk64 pk(float x, float y) { pk64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void mulw(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
    unsigned long long p; asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b)); asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
}
template <int R> __device__ __forceinline__ uint4 philox(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
        uint32_t lo0, hi0, lo1, hi1; mulw(0xD2511F53u, c.x, lo0, hi0); mulw(0xCD9E8D57u, c.z, lo1, hi1);
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float2 gauss(uint32_t w0, uint32_t w1) {
    const float u = fmaf(__uint2float_rn(w0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float lg, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
    const float t = lg * -1.3862943611198906f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    const float th = __int2float_rn((int32_t)w1) * 1.4629180792671596e-09f;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    return make_float2(r * c, r * s);
}
// MODE 0: NF FFMA2 with all-vector operands (acc = x*h + acc, h broadcast from a vector register)
// MODE 1: philox-10 calls only (NP per step); MODE 2: philox + box-muller; MODE 3: MODE 2 + NF FFMA2 (vector form)
// MODE 4: box-muller only on changing words; MODE 5: lop3 2-reg+imm with 8 FFMA2; MODE 6: lop3 3-reg with FFMA2 vector form
template <int MODE, int NF, int ROUNDS>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, unsigned key) {
    float2 acc[8], x[2]; float h[4];
    uint32_t z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f, i); z[i] = threadIdx.x * 2654435761u + i; }
    x[0] = make_float2(a, b); x[1] = make_float2(b, a); h[0] = a; h[1] = b; h[2] = a * b; h[3] = a + b;
    if (threadIdx.x == 1000) { h[0] = 3; x[0].x = 5; }   // defeat uniformity analysis: operands live in vector registers
    float2 gs = make_float2(0, 0);
    uint4 ctr = make_uint4(threadIdx.x, blockIdx.x, 0, 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0 || MODE == 3 || MODE == 6 || MODE == 5) {
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                    pk64 d = pk(acc[i & 7]);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(pk(x[i & 1])), "l"(pk(h[i & 3], h[i & 3])));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i & 7].x), "=f"(acc[i & 7].y) : "l"(d));
                    if (MODE == 5) asm volatile("lop3.b32 %0, %0, %1, 0x5555, 0x96;" : "+r"(z[i & 7]) : "r"(z[(i + 1) & 7]));
                    if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i & 7]) : "r"(z[(i + 1) & 7]), "r"(z[(i + 2) & 7]));
                }
            }
            if (MODE == 1 || MODE == 2 || MODE == 3) {
                ctr.z = it * 8 + r;
                const uint4 p = philox<ROUNDS>(ctr, key, key ^ 0x1234567u);
                if (MODE == 1) { z[0] ^= p.x ^ p.y ^ p.z ^ p.w; }
                else { const float2 g0 = gauss(p.x, p.y), g1 = gauss(p.z, p.w); gs.x += g0.x + g1.x; gs.y += g0.y + g1.y; }
            }
            if (MODE == 4) {
                z[0] += 0x9E3779B9u; z[1] += 0x7F4A7C15u;
                const float2 g0 = gauss(z[0], z[1]); gs.x += g0.x; gs.y += g0.y;
            }
        }
    }
    float s = gs.x + gs.y; unsigned zz = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += acc[i].x + acc[i].y; zz ^= z[i]; }
    if (s == 12345.678f || zz == 0x12345u) out[0] = s;
}
template <int MODE, int NF, int ROUNDS> float run(float* out, int grid, int iters, int threads = 256) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE, NF, ROUNDS><<<grid, threads>>>(out, iters, 0.999f, 1e-4f, 0x9E3779B9u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* out; cudaMalloc(&out, 64);
    const int iters = 1024;
    for (int occ = 8; occ >= 2; occ /= 2) {
        const int grid = p.multiProcessorCount * occ;   // occ CTAs of 8 warps per SM = 2*occ warps per SMSP
        const float base = run<0, 16, 10>(out, grid, iters);   // 16 FFMA2 per step -> 32 cycles if 2 cyc each
        // SMSP cycles per step: steps run by (2*occ warps per SMSP); cycles = t / t_base * 32
        auto cyc = [&](float t) { return t / base * 32.0; };
        printf("== %d warps/SMSP: 16 FFMA2 (vector operands) %.3f ms = %.1f TFLOP/s [defines 32 cyc/step]\n", 2 * occ, base, 4.0 * 16 * 8 * iters * 256.0 * grid / (base * 1e-3) / 1e12);
        printf("philox-10 call alone            : %.1f cyc\n", cyc(run<1, 0, 10>(out, grid, iters)));
        printf("philox-7 call alone             : %.1f cyc\n", cyc(run<1, 0, 7>(out, grid, iters)));
        printf("philox-10 + 2 box-muller        : %.1f cyc\n", cyc(run<2, 0, 10>(out, grid, iters)));
        printf("1 box-muller alone              : %.1f cyc\n", cyc(run<4, 0, 10>(out, grid, iters)));
        printf("philox-10 + 2 BM + 32 FFMA2 (64): %.1f cyc\n", cyc(run<3, 32, 10>(out, grid, iters)));
        printf("philox-10 + 2 BM + 64 FFMA2(128): %.1f cyc\n", cyc(run<3, 64, 10>(out, grid, iters)));
        printf("philox-10 + 2 BM + 80 FFMA2(160): %.1f cyc\n", cyc(run<3, 80, 10>(out, grid, iters)));
        printf("16 FFMA2 + 16 LOP3(2 reg + imm) : %.1f cyc\n", cyc(run<5, 16, 10>(out, grid, iters)));
        printf("16 FFMA2 + 16 LOP3(3 reg)       : %.1f cyc\n", cyc(run<6, 16, 10>(out, grid, iters)));
    }
    return 0;
}
