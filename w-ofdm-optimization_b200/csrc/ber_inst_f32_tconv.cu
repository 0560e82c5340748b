// fp32 instantiations of K1 with the channel convolution on the tensor cores (ber_tconv.cuh): N = 256, one CTA of 256
// threads per frame, two CTAs per SM; N = 512, one CTA of 512 threads per frame and SM; NTILE = tiles of 512 stream samples
// (tensor-memory accumulators) per frame.
#include "ber_registry.h"
#include "ber_tconv.cuh"
namespace wofdm {
namespace {
template <int N, int NT, int NTILE, int MINB, bool V>
struct TconvVariantImpl {
    static cudaError_t launch(const BerParams& prm, int grid, size_t smem, cudaStream_t st) {
        ber_tconv_kernel<N, NT, NTILE, MINB, V><<<grid, NT, smem, st>>>(prm);
        return cudaGetLastError();
    }
    static BerVariant make(const char* name) {
        BerVariant v;
        v.name = name; v.N = N; v.NT = NT; v.TC = 2 * NTILE; v.LB = TCV_LB; v.MINB = MINB; v.CL = 1; v.circ = false; v.txs = true; v.full = false;
        v.ntile = NTILE; v.gen = 1; v.launch_threads = NT;
        v.fp64 = false; v.verify = V;
        v.layout = &tconv_smem_layout<N, NT, NTILE>;
        v.fn = reinterpret_cast<const void*>(&ber_tconv_kernel<N, NT, NTILE, MINB, V>);
        v.launch = &launch;
        return v;
    }
};
}  // namespace
#define WOFDM_VARIANT_TCONV(N, NT, NTILE, MINB)                                                              \
    out.push_back(TconvVariantImpl<N, NT, NTILE, MINB, false>::make("ber_f32t_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB)); \
    out.push_back(TconvVariantImpl<N, NT, NTILE, MINB, true>::make("ber_f32t_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB "_verify"));
void register_ber_f32_tconv(std::vector<BerVariant>& out) {
    WOFDM_VARIANT_TCONV(256, 256, 9, 2)
    WOFDM_VARIANT_TCONV(256, 256, 10, 2)
    // N = 512: one CTA of 512 threads per frame and SM, all 512 tensor-memory columns
    WOFDM_VARIANT_TCONV(512, 512, 18, 1)
    WOFDM_VARIANT_TCONV(512, 512, 19, 1)
}
}  // namespace wofdm
