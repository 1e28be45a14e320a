// stream_pass.cu -- instantiations and launch of the warp-streaming pass kernel
// (stream_kernel.cuh); kept in its own translation unit so that the chain kernels compile
// in parallel with engine.cu.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "stream_kernel.cuh"
#include "stream_launch.h"

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

// Scheduling variant of a pass: AVIRB200_STREAM_VARIANT_H / _V (or AVIRB200_STREAM_VARIANT for
// both) override the defaults; tuning and test switch, every variant computes the same bits.
int stream_variant(bool is_v) {
    static const int v[2] = {
        [] {
            const char* e = getenv("AVIRB200_STREAM_VARIANT_H");
            if (!e) e = getenv("AVIRB200_STREAM_VARIANT");
            const int x = e ? atoi(e) : kStreamDefaultVariantH;
            return (x >= 0 && x < kStreamVariants) ? x : kStreamDefaultVariantH;
        }(),
        [] {
            const char* e = getenv("AVIRB200_STREAM_VARIANT_V");
            if (!e) e = getenv("AVIRB200_STREAM_VARIANT");
            const int x = e ? atoi(e) : kStreamDefaultVariantV;
            return (x >= 0 && x < kStreamVariants) ? x : kStreamDefaultVariantV;
        }()};
    return v[is_v ? 1 : 0];
}

int stream_launch(int chain, bool is_v, bool plain_f32, const StreamParams& p, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = -2;
    const bool known = stream_dispatch(chain, is_v, stream_variant(is_v), p.src_type, [&](auto tag, auto pass) {
        using C = typename decltype(tag)::type;
        if constexpr (!decltype(pass)::is_v) rc = launch_one<C, false, 0>(p, st);
        else rc = plain_f32 ? launch_one<C, true, 1>(p, st) : launch_one<C, true, 0>(p, st);
    });
    return known ? rc : -2;
}

} // namespace avs
