// hankel_conv.cu -- development probe (sm_100a): the L-tap complex channel convolution of a frame stream on the
// 5th-generation tensor cores, with the stream itself as the A operand.
//
// The stream is stored split, x = hi + lo (two fp16 numbers each for re and im), as half2 words hi[i], lo[i].  Row r of
// the A operand is the 24 samples 4r-20 .. 4r+3 (K = 48 halves): rows are 16 bytes apart, which is exactly the pitch of
// the rows inside an 8 x 16-byte core matrix of the K-major no-swizzle UMMA layout, so the raw array IS a canonical
// operand of the (overlapping) Hankel matrix with LBO = 16 B (next K chunk = next 4 samples), SBO = 128 B (next 8 rows).
// B (16 x 48, K-major) holds the taps: output column 2o / 2o+1 = Re / Im of y[4r + o], o < 4 (columns 8..15 are zero).
// y = hi*Thi + hi*Tlo + lo*Thi, fp32 accumulation in tensor memory; one 128 x 16 x 16 MMA per (tile of 512 samples, K step
// of 8 samples, split term).  Checks the result against a double-precision direct form and times the step.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int LB = 21, PAD = 20, NTILE = 9, LEN = NTILE * 512, SLACK = 24;
constexpr uint32_t IDESC = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // D = f32, A = B = f16, K-major, N = 16, M = 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) __trap();
}

// mode 0: one thread issues every MMA; mode 1: warp w issues tiles w, w + 8
__global__ void __launch_bounds__(256, 2) conv_kernel(const float2* __restrict__ x, const float2* __restrict__ h, float2* __restrict__ y,
                                                      int iters, int mode, long long* cycles) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    uint32_t* hi = reinterpret_cast<uint32_t*>(sm);
    uint32_t* lo = hi + (PAD + LEN + SLACK);
    unsigned char* Bh = reinterpret_cast<unsigned char*>(lo + (PAD + LEN + SLACK));
    unsigned char* Bl = Bh + 1536;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(mode == 0 ? 1 : 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // split stream
    for (int i = tid; i < PAD + LEN + SLACK; i += 256) {
        float2 v = make_float2(0.f, 0.f);
        if (i >= PAD && i < PAD + LEN) v = x[i - PAD];
        const __half2 a = __floats2half2_rn(v.x, v.y);
        const float2 af = __half22float2(a);
        const __half2 b = __floats2half2_rn(v.x - af.x, v.y - af.y);
        hi[i] = *reinterpret_cast<const uint32_t*>(&a);
        lo[i] = *reinterpret_cast<const uint32_t*>(&b);
    }
    // taps matrix: n = 2o + comp, k = 2j + d, l = 20 + o - j
    for (int e = tid; e < 16 * 24; e += 256) {
        const int n = e / 24, j = e % 24;
        const int o = n >> 1, comp = n & 1, l = 20 + o - j;
        float2 t = make_float2(0.f, 0.f);
        if (n < 8 && l >= 0 && l < LB) t = h[l];
        const float2 v = comp == 0 ? make_float2(t.x, -t.y) : make_float2(t.y, t.x);   // (d = 0, d = 1)
        const __half2 a = __floats2half2_rn(v.x, v.y);
        const float2 af = __half22float2(a);
        const __half2 b = __floats2half2_rn(v.x - af.x, v.y - af.y);
        // element (n, k = 2j..2j+1): n-group G = n/8, K chunk c = j/4, word j%4 of the 16-byte row
        const int off = (n >> 3) * 768 + (j >> 2) * 128 + (n & 7) * 16 + (j & 3) * 4;
        *reinterpret_cast<uint32_t*>(Bh + off) = *reinterpret_cast<const uint32_t*>(&a);
        *reinterpret_cast<uint32_t*>(Bl + off) = *reinterpret_cast<const uint32_t*>(&b);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_hi = smem_u32(hi), a_lo = smem_u32(lo), b_hi = smem_u32(Bh), b_lo = smem_u32(Bl);

    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        auto issue_tile = [&](int t) {
            const uint32_t tacc = tmem + (uint32_t)(16 * t);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint64_t dah = umma_desc(a_hi + t * 2048 + j * 32, 16, 128), dal = umma_desc(a_lo + t * 2048 + j * 32, 16, 128);
                const uint64_t dbh = umma_desc(b_hi + j * 256, 128, 768), dbl = umma_desc(b_lo + j * 256, 128, 768);
                mma_f16(tacc, dah, dbh, j != 0);
                mma_f16(tacc, dah, dbl, 1u);
                mma_f16(tacc, dal, dbh, 1u);
            }
        };
        if (mode == 0) {
            if (tid == 0) {
                for (int t = 0; t < NTILE; ++t) issue_tile(t);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            }
        } else {
            if (lane == 0) {
                for (int t = warp; t < NTILE; t += 8) issue_tile(t);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            }
        }
        mbar_wait(smem_u32(&bar), (uint32_t)(it & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int wg = warp >> 2, row = (warp & 3) * 32 + lane;
        float2 accsum = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
            uint32_t v0, v1, v2, v3;
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(16 * t + 4 * wg);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int p = 512 * t + 4 * row + 2 * wg;
            if (it == iters - 1) {
                y[p] = make_float2(__uint_as_float(v0), __uint_as_float(v1));
                y[p + 1] = make_float2(__uint_as_float(v2), __uint_as_float(v3));
            } else {
                accsum.x += __uint_as_float(v0) + __uint_as_float(v2);
                accsum.y += __uint_as_float(v1) + __uint_as_float(v3);
            }
        }
        if (accsum.x == 1234.5f) y[0] = accsum;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}

int main() {
    std::vector<float2> x(LEN), h(LB), y(LEN);
    srand(1);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& v : x) v = make_float2(16.f * rnd(), 16.f * rnd());
    for (int l = 0; l < LB; ++l) { const float s = expf(-0.25f * l); h[l] = make_float2(s * rnd(), s * rnd()); }
    float2 *dx, *dh, *dy; long long* dc;
    cudaMalloc(&dx, LEN * 8); cudaMalloc(&dh, LB * 8); cudaMalloc(&dy, LEN * 8); cudaMalloc(&dc, 8);
    cudaMemcpy(dx, x.data(), LEN * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dh, h.data(), LB * 8, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(PAD + LEN + SLACK) * 8 + 2 * 1536;
    cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<double> ref(2 * LEN);
    double rms = 0;
    for (int i = 0; i < LEN; ++i) {
        double re = 0, im = 0;
        for (int l = 0; l < LB && l <= i; ++l) {
            re += (double)h[l].x * x[i - l].x - (double)h[l].y * x[i - l].y;
            im += (double)h[l].x * x[i - l].y + (double)h[l].y * x[i - l].x;
        }
        ref[2 * i] = re; ref[2 * i + 1] = im; rms += re * re + im * im;
    }
    rms = sqrt(rms / LEN);
    for (int mode = 0; mode < 2; ++mode) {
        for (int iters : {1, 200}) {
            cudaMemset(dy, 0, LEN * 8);
            conv_kernel<<<iters == 1 ? 1 : 296, 256, smem>>>(dx, dh, dy, iters, mode, dc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d iters %d: CUDA error %s\n", mode, iters, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(y.data(), dy, LEN * 8, cudaMemcpyDeviceToHost);
            long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
            double emax = 0; int imax = 0;
            for (int i = 0; i < LEN; ++i) {
                const double d = fmax(fabs(y[i].x - ref[2 * i]), fabs(y[i].y - ref[2 * i + 1]));
                if (d > emax) { emax = d; imax = i; }
            }
            printf("mode %d iters %3d: max abs err %.3e (rms %.3e, rel %.3e) at %d; cycles/iter %.0f\n", mode, iters, emax, rms, emax / rms, imax,
                   (double)cyc / iters);
        }
    }
    printf("y[0..2] = (%f,%f) (%f,%f) ref (%f,%f) (%f,%f)\n", y[0].x, y[0].y, y[1].x, y[1].y, ref[0], ref[1], ref[2], ref[3]);
    return 0;
}
