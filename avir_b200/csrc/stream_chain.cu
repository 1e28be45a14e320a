// stream_chain.cu -- the kernels of ONE pass of ONE chain of the warp-streaming pass kernel
// (stream_kernel.cuh) and their launch.  Compiled once per chain and pass with
// -DAVS_CHAIN_ID=<id> -DAVS_CHAIN_PASS=<0|1> (avir_b200/build.py), so that the units build in
// parallel; stream_pass.cu routes a launch to the right one.
#include <cuda.h>
#include <cuda_runtime.h>

#include "stream_kernel.cuh"
#include "stream_launch.h"

#if !defined(AVS_CHAIN_ID) || !defined(AVS_CHAIN_PASS)
#error "compile with -DAVS_CHAIN_ID=<StreamChainId> -DAVS_CHAIN_PASS=<0 row pass | 1 column pass>"
#endif

namespace avs {

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static const EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// The column pass's source as a 2-D fp32 tensor: rows of n_lines pixels x 4 channels from p.src
// on (row 0 = global intermediate row src_row_base), box = one ring group of a warp.
bool encode_source_tensor(StreamParams& p, int box_rows) {
    static_assert(sizeof(CUtensorMap) == sizeof(p.tmap), "tensor map size");
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)p.n_lines * 4, (cuuint64_t)(p.src_len - p.src_row_base)};
    const cuuint64_t strides[1] = {(cuuint64_t)p.src_pitch * 4};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (dims[1] < 1) return false;
    return enc(reinterpret_cast<CUtensorMap*>(p.tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(p.src),
               dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class C, bool IS_V, int EPI>
int launch_one(const StreamParams& p_in, int sm_count, cudaStream_t st) {
    StreamParams p = p_in;
    if constexpr (C::MBAR) {
        if (!encode_source_tensor(p, C::SRC_N)) return -3; // no tensor map: the caller falls back to variant 1
    }
    constexpr int NW = IS_V ? C::NWARPS_V : C::NWARPS_H;
    constexpr size_t smem = (size_t)NW * (IS_V ? C::WARP_F2_V : C::WARP_F2_H) * sizeof(float2);
    static_assert(smem <= 227 * 1024, "per-warp rings do not fit the shared memory of an SM");
    auto kern = stream_pass_kernel<C, IS_V, EPI>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    // one persistent block per SM; fewer when the pass has fewer rounds than warps
    const long long rps = (long long)(p.out1 - 1) / C::B - p.out0 / C::B + 1;
    const long long units = rps * stream_strip_count(p);
    long long blocks = (units + NW - 1) / NW;
    if (blocks > sm_count) blocks = sm_count;
    if (blocks < 1) return 0;
    kern<<<(int)blocks, NW * 32, smem, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

} // namespace

template <>
int stream_launch_chain<AVS_CHAIN_ID, (AVS_CHAIN_PASS != 0)>(int variant, int epi, const StreamParams& p, int sm_count,
                                                             void* stream) {
    constexpr bool IS_V = (AVS_CHAIN_PASS != 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = -2;
    auto run = [&](int var) {
        return stream_dispatch_chain<AVS_CHAIN_ID>(IS_V, var, p.src_type, [&](auto tag, auto pass) {
            using C = typename decltype(tag)::type;
            if constexpr (decltype(pass)::is_v != IS_V) {
                (void)rc; // (the dispatcher instantiates the callback for both passes)
            } else if constexpr (!IS_V) {
                rc = (p.xs_count != nullptr) ? launch_one<C, false, kEpiXs>(p, sm_count, st) // sender of the fused halo exchange
                                             : launch_one<C, false, 0>(p, sm_count, st);
            } else {
                rc = (epi == 1) ? launch_one<C, true, 1>(p, sm_count, st)
                                : (epi == 2 ? launch_one<C, true, 2>(p, sm_count, st) : launch_one<C, true, 0>(p, sm_count, st));
            }
        });
    };
    bool known = run(variant);
    if (known && rc == -3) known = run(1); // tensor map could not be encoded (old driver): per-lane copies
    return known ? rc : -2;
}

} // namespace avs
