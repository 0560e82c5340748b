// ber_registry.h -- host-side table of the compiled K1 variants.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "ber_kernel.cuh"

namespace wofdm {

struct BerVariant {
    const char* name;
    int N, NT, TC, LB, MINB;
    bool full;                 // every register of every thread is a live stream sample (NT*TC == S*stride)
    bool fp64, verify;
    BerSmem (*layout)(int S, int stride, int tail_tx, int tail_rx, int L, int chunk, int use_global);
    const void* fn;
    cudaError_t (*launch)(const BerParams& prm, int grid, size_t smem, cudaStream_t st);
};

template <typename T, int N, int NT, int TC, int LB, int MINB, bool FULL, bool V>
struct BerVariantImpl {
    static cudaError_t launch(const BerParams& prm, int grid, size_t smem, cudaStream_t st) {
        ber_frame_kernel<T, N, NT, TC, LB, MINB, FULL, V><<<grid, NT, smem, st>>>(prm);
        return cudaGetLastError();
    }
    static BerVariant make(const char* name) {
        BerVariant v;
        v.name = name; v.N = N; v.NT = NT; v.TC = TC; v.LB = LB; v.MINB = MINB; v.full = FULL;
        v.fp64 = sizeof(T) == 8; v.verify = V;
        v.layout = &ber_smem_layout<T, N, NT, TC, LB>;
        v.fn = reinterpret_cast<const void*>(&ber_frame_kernel<T, N, NT, TC, LB, MINB, FULL, V>);
        v.launch = &launch;
        return v;
    }
};

#define WOFDM_VARIANT(T, N, NT, TC, LB, MINB, FULL, tag)                                            \
    out.push_back(BerVariantImpl<T, N, NT, TC, LB, MINB, FULL, false>::make("ber_" tag "_n" #N "_t" #NT "_c" #TC "_l" #LB "_b" #MINB "_f" #FULL)); \
    out.push_back(BerVariantImpl<T, N, NT, TC, LB, MINB, FULL, true>::make("ber_" tag "_n" #N "_t" #NT "_c" #TC "_l" #LB "_b" #MINB "_f" #FULL "_verify"));

void register_ber_f32_staged(std::vector<BerVariant>& out);
void register_ber_f64_staged(std::vector<BerVariant>& out);
void register_ber_f32_regs(std::vector<BerVariant>& out);

}  // namespace wofdm
