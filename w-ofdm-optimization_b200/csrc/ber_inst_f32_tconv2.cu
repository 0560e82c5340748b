// fp32 instantiations of K1, second generation of the tensor-core-convolution kernel (ber_tconv2.cuh): N = 256, one CTA
// of 256 threads per frame, two CTAs per SM; N = 512, one CTA of 512 threads per frame and SM; NTILE = tiles of 512 stream
// samples (tensor-memory accumulators) per frame.
#include "ber_registry.h"
#include "ber_tconv2.cuh"
namespace wofdm {
namespace {
template <int N, int NT, int NTILE, int MINB, bool V, int CL = 1, int LB = TCV_LB, bool TXY = false>
struct Tconv2VariantImpl {
    // grid = CTAs (a multiple of CL); CL > 1 launches thread-block clusters of CL CTAs
    static cudaError_t launch(const BerParams& prm, int grid, size_t smem, cudaStream_t st) {
        if constexpr (CL == 1) {
            ber_tconv2_kernel<N, NT, NTILE, MINB, V, 1, LB, TXY><<<grid, NT + tconv2_mma_warp_threads(N, NT), smem, st>>>(prm);   // (+ the MMA warp)
            return cudaGetLastError();
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT + tconv2_mma_warp_threads(N, NT)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            return cudaLaunchKernelEx(&cfg, ber_tconv2_kernel<N, NT, NTILE, MINB, V, CL, LB, TXY>, prm);
        }
    }
    static BerVariant make(const char* name) {
        BerVariant v;
        v.name = name; v.N = N; v.NT = NT; v.TC = 2 * NTILE; v.LB = LB; v.MINB = MINB; v.CL = CL; v.circ = false; v.txs = true; v.full = false;
        v.ntile = NTILE; v.gen = 2; v.txy = TXY; v.launch_threads = NT + tconv2_mma_warp_threads(N, NT);
        v.fp64 = false; v.verify = V;
        v.layout = &tconv2_smem_layout<N, NT, NTILE, LB>;
        v.fn = reinterpret_cast<const void*>(&ber_tconv2_kernel<N, NT, NTILE, MINB, V, CL, LB, TXY>);
        v.launch = &launch;
        return v;
    }
};
}  // namespace
#define WOFDM_VARIANT_TCONV2(N, NT, NTILE, MINB)                                                              \
    out.push_back(Tconv2VariantImpl<N, NT, NTILE, MINB, false>::make("ber_f32t2_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB)); \
    out.push_back(Tconv2VariantImpl<N, NT, NTILE, MINB, true>::make("ber_f32t2_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB "_verify"));
// CL CTAs per frame (thread-block cluster): NTILE tiles per CTA
#define WOFDM_VARIANT_TCONV2_CL(N, NT, NTILE, MINB, CL)                                                       \
    out.push_back(Tconv2VariantImpl<N, NT, NTILE, MINB, false, CL>::make("ber_f32t2_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB "_cl" #CL)); \
    out.push_back(Tconv2VariantImpl<N, NT, NTILE, MINB, true, CL>::make("ber_f32t2_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB "_cl" #CL "_verify"));
// long channels: LB taps of convolution history in the Hankel operand (22 MMAs per tile at LB = 84 instead of 6)
#define WOFDM_VARIANT_TCONV2_LB(N, NT, NTILE, MINB, CL, LB)                                                   \
    out.push_back(Tconv2VariantImpl<N, NT, NTILE, MINB, false, CL, LB>::make("ber_f32t2_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB "_cl" #CL "_l" #LB)); \
    out.push_back(Tconv2VariantImpl<N, NT, NTILE, MINB, true, CL, LB>::make("ber_f32t2_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB "_cl" #CL "_l" #LB "_verify"));
// channel-mask chain: the Tx stream gathered from the mask product's output (BerParams::tx_y, tconv2_load_masked)
#define WOFDM_VARIANT_TCONV2_TXY(N, NT, NTILE, MINB)                                                          \
    out.push_back(Tconv2VariantImpl<N, NT, NTILE, MINB, false, 1, TCV_LB, true>::make("ber_f32t2_n" #N "_t" #NT "_tiles" #NTILE "_b" #MINB "_txy"));
void register_ber_f32_tconv2(std::vector<BerVariant>& out) {
    WOFDM_VARIANT_TCONV2_TXY(256, 256, 9, 2)
    WOFDM_VARIANT_TCONV2_TXY(256, 256, 10, 2)
    WOFDM_VARIANT_TCONV2_TXY(512, 512, 18, 1)
    WOFDM_VARIANT_TCONV2_TXY(512, 512, 19, 1)
    WOFDM_VARIANT_TCONV2(256, 256, 9, 2)
    WOFDM_VARIANT_TCONV2(256, 256, 10, 2)
    WOFDM_VARIANT_TCONV2(512, 512, 18, 1)
    WOFDM_VARIANT_TCONV2(512, 512, 19, 1)
    // N = 1024: a cluster of two CTAs of 512 threads per frame, 8 OFDM symbols and 17..19 tiles per CTA, one CTA per SM
    WOFDM_VARIANT_TCONV2_CL(1024, 512, 18, 1, 2)
    WOFDM_VARIANT_TCONV2_CL(1024, 512, 19, 1, 2)
    // up to 84 taps (BASELINE configs[4]: "L = 21, optionally 84"), N = 256 and the N = 1024 cluster kernel
    WOFDM_VARIANT_TCONV2_LB(256, 256, 9, 2, 1, 84)
    WOFDM_VARIANT_TCONV2_LB(256, 256, 10, 2, 1, 84)
    WOFDM_VARIANT_TCONV2_LB(1024, 512, 18, 1, 2, 84)
    WOFDM_VARIANT_TCONV2_LB(1024, 512, 19, 1, 2, 84)
    // (measured and not kept: the same frames on clusters of 256-thread CTAs, two CTAs of different frames per SM --
    //  N = 512 as 2 x 256 threads: 1.98e8 OFDM symbols/s against 2.06e8; N = 1024 as 4 x 256: 5.5e7 against 8.4e7)
}
}  // namespace wofdm
