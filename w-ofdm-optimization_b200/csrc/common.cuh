// common.cuh -- small device helpers shared by the w-OFDM kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wofdm {

template <typename T> struct V2sel;
template <> struct V2sel<float> { typedef float2 type; };
template <> struct V2sel<double> { typedef double2 type; };
template <typename T> using V2 = typename V2sel<T>::type;

template <typename T> __host__ __device__ __forceinline__ V2<T> mk2(T x, T y) { V2<T> r; r.x = x; r.y = y; return r; }
// ---------------------------------------------------------------------------------------------
// Complex arithmetic.  double2: plain scalar code.  float2: Blackwell packed-FP32 instructions
// (add/mul/fma.f32x2 -> SASS FADD2/FMUL2/FFMA2): one issue slot per complex add, two per complex
// multiply-accumulate.  The "(-a.y, a.x)" operands below cost nothing: ptxas folds the half swap
// and the sign into the FFMA2/FADD2 operand modifiers (.F32x2.LO_HI.NP), and a duplicated scalar
// into the broadcast form (.F32).
// ---------------------------------------------------------------------------------------------
typedef unsigned long long pk64;
__device__ __forceinline__ pk64 pk(float x, float y) { pk64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ pk64 pk(float2 a) { return pk(a.x, a.y); }
__device__ __forceinline__ float2 upk(pk64 r) { float2 d; asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r)); return d; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { pk64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { pk64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { pk64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { pk64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return upk(d); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return sub2(a, b); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { a.x += b.x; a.y += b.y; return a; }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { a.x -= b.x; a.y -= b.y; return a; }
// i*a and -i*a
template <typename V> __device__ __forceinline__ V rotj(V a) { V r; r.x = -a.y; r.y = a.x; return r; }
template <typename V> __device__ __forceinline__ V rotmj(V a) { V r; r.x = a.y; r.y = -a.x; return r; }
// real scalar times complex
__device__ __forceinline__ float2 cscale(float s, float2 a) { return mul2(a, make_float2(s, s)); }
__device__ __forceinline__ double2 cscale(double s, double2 a) { a.x *= s; a.y *= s; return a; }
// c + s*a (s real)
__device__ __forceinline__ float2 caxpy(float s, float2 a, float2 c) { return fma2(a, make_float2(s, s), c); }
__device__ __forceinline__ double2 caxpy(double s, double2 a, double2 c) { c.x = fma(s, a.x, c.x); c.y = fma(s, a.y, c.y); return c; }
// a*b = b.x*a + b.y*(i a)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return caxpy(b.y, rotj(a), cscale(b.x, a)); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { double2 r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
// a*conj(b) = b.x*a + b.y*(-i a)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return caxpy(b.y, rotmj(a), cscale(b.x, a)); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { double2 r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r; }
// acc += h*x
__device__ __forceinline__ void cmac(float2& acc, float2 h, float2 x) { acc = caxpy(h.x, x, acc); acc = caxpy(h.y, rotj(x), acc); }
__device__ __forceinline__ void cmac(double2& acc, double2 h, double2 x) {
    acc.x = fma(h.x, x.x, acc.x); acc.x = fma(-h.y, x.y, acc.x);
    acc.y = fma(h.x, x.y, acc.y); acc.y = fma(h.y, x.x, acc.y);
}
// p += (a.x^2, a.y^2): running |a|^2 in two lanes, summed by the caller
__device__ __forceinline__ float2 csq_acc(float2 a, float2 p) { return fma2(a, a, p); }
__device__ __forceinline__ double2 csq_acc(double2 a, double2 p) { p.x = fma(a.x, a.x, p.x); p.y = fma(a.y, a.y, p.y); return p; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter-based: draws depend only on (key, counter).
// ---------------------------------------------------------------------------------------------
// One round = two 32x32->64 products (one IMAD.WIDE.U32 each on the device; the halves are taken with
// mov.b64 so that no 64-bit shift is left for the compiler to lower) and two 3-input XORs.
__host__ __device__ __forceinline__ void mul_wide32(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
#ifdef __CUDA_ARCH__
    unsigned long long p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
#else
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
#endif
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t lo0, hi0, lo1, hi1;
        mul_wide32(0xD2511F53u, c.x, lo0, hi0);
        mul_wide32(0xCD9E8D57u, c.z, lo1, hi1);
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// The same function with the ten round keys precomputed (they depend on the seed only): taken from a kernel parameter
// they are constant-bank operands of the XORs, and the twenty key additions per call disappear.
__host__ __device__ inline void philox_round_keys(unsigned long long seed, uint32_t (&rk)[20]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) { rk[2 * i] = k0; rk[2 * i + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
__host__ __device__ __forceinline__ uint4 philox4x32_10_rk(uint4 c, const uint32_t (&rk)[20]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t lo0, hi0, lo1, hi1;
        mul_wide32(0xD2511F53u, c.x, lo0, hi0);
        mul_wide32(0xCD9E8D57u, c.z, lo1, hi1);
        c = make_uint4(hi1 ^ c.y ^ rk[2 * i], lo1, hi0 ^ c.w ^ rk[2 * i + 1], lo0);
    }
    return c;
}

// Counter layout: (frame_lo, frame_hi, index, stream).
//   stream 0            : constellation indices, index = s*(N/16) + t  -> 16 bytes = sub-carriers
//                         t + q*(N/16), q = 0..15 of OFDM symbol s (low `bits` bits of byte q)
//   stream 1 + variant  : noise, index = q -> complex samples 2q (words x,y) and 2q+1 (words z,w)
enum { STREAM_SYM = 0, STREAM_NOISE = 1 };

// Box-Muller: one complex sample with independent N(0,1) parts from two 32-bit words.
__device__ __forceinline__ float2 gauss_pair(uint32_t w0, uint32_t w1, float) {
    // u in (0,1]: (w0 + 0.5) * 2^-32, exact enough in fp32; theta in [-pi, pi)
    const float u = fmaf(__uint2float_rn(w0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float lg, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
    const float t = lg * -1.3862943611198906f;               // -2 ln u >= 0 (u <= 1 after rounding; -0 is harmless)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    const float th = __int2float_rn((int32_t)w1) * 1.4629180792671596e-09f;  // 2*pi * 2^-32
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    return make_float2(r * c, r * s);
}
// Two complex samples from one Philox call (words x,y -> n0; z,w -> n1): the same arithmetic as gauss_pair, with the
// scalar FP32 steps of the two samples paired into packed instructions (a scalar FFMA/FMUL next to FFMA2s costs
// almost as much dispatch time as a packed one).  Bit-identical to two gauss_pair calls.
__device__ __forceinline__ void gauss_quad(uint4 r, float2& n0, float2& n1, float) {
    const float2 u = fma2(make_float2(__uint2float_rn(r.x), __uint2float_rn(r.z)),
                          make_float2(2.3283064365386963e-10f, 2.3283064365386963e-10f),
                          make_float2(1.1641532182693481e-10f, 1.1641532182693481e-10f));
    float2 lg, rr;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg.x) : "f"(u.x));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg.y) : "f"(u.y));
    const float2 t = mul2(lg, make_float2(-1.3862943611198906f, -1.3862943611198906f));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rr.x) : "f"(t.x));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rr.y) : "f"(t.y));
    const float2 th = mul2(make_float2(__int2float_rn((int32_t)r.y), __int2float_rn((int32_t)r.w)),
                           make_float2(1.4629180792671596e-09f, 1.4629180792671596e-09f));
    float s0, c0, s1, c1;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(th.x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(th.x));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(th.y));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(th.y));
    n0 = mul2(make_float2(c0, s0), make_float2(rr.x, rr.x));
    n1 = mul2(make_float2(c1, s1), make_float2(rr.y, rr.y));
}
__device__ __forceinline__ double2 gauss_pair(uint32_t w0, uint32_t w1, double);
__device__ __forceinline__ void gauss_quad(uint4 r, double2& n0, double2& n1, double) {
    n0 = gauss_pair(r.x, r.y, double());
    n1 = gauss_pair(r.z, r.w, double());
}
__device__ __forceinline__ double2 gauss_pair(uint32_t w0, uint32_t w1, double) {
    const double u = ((double)w0 + 0.5) * 2.3283064365386963e-10;
    const double r = sqrt(-2.0 * log(u));
    double s, c;
    sincospi((double)(int32_t)w1 * 4.656612873077393e-10, &s, &c);  // 2 * 2^-32
    return make_double2(r * c, r * s);
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace wofdm
