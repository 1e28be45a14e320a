// stream_pass.cu -- launch entry of the warp-streaming pass kernel: picks the scheduling variant
// and routes to the chain's own translation unit (stream_chain.cu, compiled once per chain).
#include <stdlib.h>

#include "stream_launch.h"

namespace avs {

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

int stream_launch(int chain, bool is_v, int epi, const StreamParams& p, void* stream) {
    const int var = stream_variant(is_v);
    switch (chain) {
    case kChainDil24: return stream_launch_chain<kChainDil24>(is_v, var, epi, p, stream);
    case kChainInl24: return stream_launch_chain<kChainInl24>(is_v, var, epi, p, stream);
    case kChainInl3: return stream_launch_chain<kChainInl3>(is_v, var, epi, p, stream);
    case kChainInl3D: return stream_launch_chain<kChainInl3D>(is_v, var, epi, p, stream);
#ifdef AVS_WITH_DIL56
    // Not in the default build: the chain loses to the tile kernel on B200 (stream_types.h) and its
    // kernels take ~18 minutes to compile; AVIRB200_BUILD_ALL_CHAINS=1 python avir_b200/build.py
    // builds them (then selectable with AVIRB200_STREAM_ALL=1).  Without them the engine falls
    // back to the tile kernel (return -2 = no such instantiation).
    case kChainDil56: return stream_launch_chain<kChainDil56>(is_v, var, epi, p, stream);
#endif
    case kChainUp2: return stream_launch_chain<kChainUp2>(is_v, var, epi, p, stream);
    default: return -2;
    }
}

} // namespace avs
