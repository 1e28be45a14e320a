// stream_chain.cu -- the kernels of ONE chain of the warp-streaming pass kernel
// (stream_kernel.cuh) and their launch.  Compiled once per chain with -DAVS_CHAIN_ID=<id>
// (avir_b200/build.py), so that the chains build in parallel; stream_pass.cu routes a launch
// to the right one.
#include <cuda_runtime.h>

#include "stream_kernel.cuh"
#include "stream_launch.h"

#ifndef AVS_CHAIN_ID
#error "compile with -DAVS_CHAIN_ID=<StreamChainId>"
#endif

namespace avs {

namespace {

int sm_count() {
    static const int sms = [] {
        int d = 0, n = 0;
        cudaGetDevice(&d);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
        return n > 0 ? n : 1;
    }();
    return sms;
}

template <class C, bool IS_V, int EPI>
int launch_one(const StreamParams& p, cudaStream_t st) {
    constexpr int NW = IS_V ? C::NWARPS_V : C::NWARPS_H;
    constexpr size_t smem = (size_t)NW * (IS_V ? C::WARP_F2_V : C::WARP_F2_H) * sizeof(float2);
    static_assert(smem <= 227 * 1024, "per-warp rings do not fit the shared memory of an SM");
    auto kern = stream_pass_kernel<C, IS_V, EPI>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    // one persistent block per SM; fewer when the pass has fewer rounds than warps
    const long long rps = (long long)(p.out1 - 1) / C::B - p.out0 / C::B + 1;
    const long long units = rps * ((p.n_lines + kLines - 1) / kLines);
    long long blocks = (units + NW - 1) / NW;
    if (blocks > sm_count()) blocks = sm_count();
    if (blocks < 1) return 0;
    kern<<<(int)blocks, NW * 32, smem, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

} // namespace

template <>
int stream_launch_chain<AVS_CHAIN_ID>(bool is_v, int variant, int epi, const StreamParams& p, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = -2;
    const bool known = stream_dispatch_chain<AVS_CHAIN_ID>(is_v, variant, p.src_type, [&](auto tag, auto pass) {
        using C = typename decltype(tag)::type;
        if constexpr (!decltype(pass)::is_v) rc = launch_one<C, false, 0>(p, st);
        else rc = (epi == 1) ? launch_one<C, true, 1>(p, st)
                             : (epi == 2 ? launch_one<C, true, 2>(p, st) : launch_one<C, true, 0>(p, st));
    });
    return known ? rc : -2;
}

} // namespace avs
