// interf_tf32.cu -- K2b', the interference contraction on 5th-generation tensor cores (tcgen05, kind::tf32)
// with a 3xTF32 split for fp32-grade accuracy (wofdm_interf_power mode 1).
//
//   [Re A; Im A] = Rbig . [Re B; Im B]       (interf.cu), one (channel, slice) = one 256-column block.
// Every fp64 operand x is split as x = hi + lo, hi = fp32(x) with the 13 low mantissa bits cleared (exactly a
// TF32 number), lo = fp32(x - hi) (the tensor core reads its top 19 bits); A.B ~ Ahi.Bhi + Ahi.Blo + Alo.Bhi,
// relative error ~2^-21 per product, accumulated in fp32 in tensor memory.
//
// One CTA (128 threads) owns a 128 x 256 accumulator tile in TMEM (256 of the 512 columns).  Per K block of 32:
// all threads copy the hi/lo tiles of both operands into shared memory in the canonical K-major no-swizzle UMMA
// layout (8-row x 16-byte core matrices; chunk c of a tile sits at byte 16*c), one elected thread issues the 12
// tcgen05.mma (4 K-steps x 3 split terms) and commits them to an mbarrier; two stages, so the copy of block k+1
// overlaps the MMAs of block k.  Epilogue: tcgen05.ld the accumulator rows, mask the diagonal of slice 0, square,
// sum, one atomicAdd per row -- A itself is never written.
#include "interf.h"

namespace wofdm {

namespace {

constexpr int TM = 128, TN = 256, KB = 32, NSTAGE = 2;
constexpr int A_BYTES = TM * KB * 4, B_BYTES = TN * KB * 4;            // one hi (or lo) tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;                 // 96 KiB
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: LBO = 128 B between the two 16-byte K chunks of one MMA, SBO = 1024 B between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((128u >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version of sm_100
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) __trap();                          // never spin forever on a GPU we share
}

// chunk c (16 bytes) of a K-major tile: row (c/64)*8 + c%8, K chunk (c%64)/8; lands at shared byte 16*c
__device__ __forceinline__ void copy_tile(unsigned char* dst, const float* __restrict__ src, int rows, int ld, int k0,
                                          int tid, int nt) {
    const int nchunk = rows * (KB / 4);
    for (int c = tid; c < nchunk; c += nt) {
        const int row = (c >> 6) * 8 + (c & 7), kc = (c & 63) >> 3;
        const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)row * ld + k0 + kc * 4);
        *reinterpret_cast<uint4*>(dst + (size_t)c * 16) = v;
    }
}

}  // namespace

// grid (2N / TM, slices), 128 threads, dynamic smem NSTAGE * STAGE_BYTES
__global__ void __launch_bounds__(128, 1) gemm_power_tf32(const float* __restrict__ Rhi, const float* __restrict__ Rlo,
                                                         const float* __restrict__ Bhi, const float* __restrict__ Blo,
                                                         double* __restrict__ P, int N, int Kp, int Ms, int c0, int scalar) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bars[NSTAGE];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TM, s = blockIdx.y;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(TN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    const float* bh = Bhi + (size_t)s * TN * Kp;
    const float* bl = Blo + (size_t)s * TN * Kp;
    const int nk = Kp / KB;
    for (int kb = 0; kb < nk; ++kb) {
        const int st = kb & 1;
        unsigned char* base = sm + (size_t)st * STAGE_BYTES;
        if (kb >= NSTAGE) mbar_wait(smem_u32(&bars[st]), (uint32_t)(((kb >> 1) - 1) & 1));   // MMAs of block kb-2 have drained this stage
        copy_tile(base, Rhi + (size_t)m0 * Kp, TM, Kp, kb * KB, tid, 128);
        copy_tile(base + A_BYTES, Rlo + (size_t)m0 * Kp, TM, Kp, kb * KB, tid, 128);
        copy_tile(base + 2 * A_BYTES, bh, TN, Kp, kb * KB, tid, 128);
        copy_tile(base + 2 * A_BYTES + B_BYTES, bl, TN, Kp, kb * KB, tid, 128);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(base), a_lo = a_hi + A_BYTES, b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) {                             // one MMA consumes 8 TF32 = two 16-byte chunks
                const uint32_t off = (uint32_t)j * 256u;
                mma_tf32(tmem, umma_desc(a_hi + off), umma_desc(b_hi + off), (kb | j) != 0);
                mma_tf32(tmem, umma_desc(a_hi + off), umma_desc(b_lo + off), 1u);
                mma_tf32(tmem, umma_desc(a_lo + off), umma_desc(b_hi + off), 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bars[st])) : "memory");
        }
    }
    // the last commit covers every MMA issued before it
    {
        const int kb = nk - 1;
        mbar_wait(smem_u32(&bars[kb & 1]), (uint32_t)((kb >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: thread = one row of the tile (TMEM lane 32*warp + lane)
    const int r = m0 + warp * 32 + lane, k = r % N;
    const int c = c0 + s / Ms, ms = s % Ms;
    float pw = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < TN / 32; ++cc) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x = __uint_as_float(v[i]);
            if (!(ms == 0 && cc * 32 + i == k)) pw = fmaf(x, x, pw);
        }
    }
    atomicAdd(scalar ? &P[c] : &P[(size_t)c * N + k], (double)pw);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TN) : "memory");
}

// x -> (hi, lo) fp32 pair, elementwise
__global__ void split_tf32(const double* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float h = __uint_as_float(__float_as_uint((float)x[i]) & 0xffffe000u);
        hi[i] = h;
        lo[i] = (float)(x[i] - (double)h);
    }
}

// [slice][kk][j] fp64 -> [slice][j][kk] hi/lo fp32 (K-major for the tensor core), 32x32 tiles through smem
__global__ void __launch_bounds__(256) transpose_split_b(const double* __restrict__ Bbig, float* __restrict__ Bhi,
                                                         float* __restrict__ Blo, int N, int Kp) {
    __shared__ double tile[32][33];
    const int s = blockIdx.z, k0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double* src = Bbig + (size_t)s * Kp * N;
    for (int i = ty; i < 32; i += 8) tile[i][tx] = src[(size_t)(k0 + i) * N + j0 + tx];
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const double x = tile[tx][i];
        const float h = __uint_as_float(__float_as_uint((float)x) & 0xffffe000u);
        const size_t o = ((size_t)s * N + j0 + i) * Kp + k0 + tx;
        Bhi[o] = h;
        Blo[o] = (float)(x - (double)h);
    }
}

int interf_gemm_tf32(wofdm_ctx* h, const wofdm_sys_t* sys, const InterfDev& v, int Ms, int c0, int slices, int scalar) {
    DeviceCtx& d = h->devs[0];
    const int N = sys->N;
    if (N != TN) return fail(h, WOFDM_EUNSUPPORTED, "TF32-split interference path is built for N = 256");
    if (v.Kp % KB) return fail(h, WOFDM_EINVAL, "Kp must be a multiple of 32");
    const size_t nr = (size_t)2 * N * v.Kp, nb = (size_t)slices * N * v.Kp;
    // fp32 work buffers live behind the fp64 B matrix of this batch (interf_upload reserved room for them)
    float* Rhi = v.tf32_work;
    float* Rlo = Rhi + nr;
    float* Bhi = Rlo + nr;
    float* Blo = Bhi + nb;
    split_tf32<<<256, 256, 0, d.stream>>>(v.Rbig, Rhi, Rlo, nr);
    transpose_split_b<<<dim3(N / 32, v.Kp / 32, slices), 256, 0, d.stream>>>(v.Bbig, Bhi, Blo, N, v.Kp);
    WOFDM_CUDA(h, cudaGetLastError());
    const size_t smem = (size_t)NSTAGE * STAGE_BYTES;
    WOFDM_CUDA(h, cudaFuncSetAttribute(gemm_power_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_power_tf32<<<dim3(2 * N / TM, slices), 128, smem, d.stream>>>(Rhi, Rlo, Bhi, Blo, v.P, N, v.Kp, Ms, c0, scalar);
    WOFDM_CUDA(h, cudaGetLastError());
    h->launches += 2;
    return WOFDM_OK;
}

}  // namespace wofdm
