// stream_pass.cu -- launch entry of the warp-streaming pass kernel: routes to the chain's and
// pass's own translation unit (stream_chain.cu, compiled once per chain and pass).
#include "stream_launch.h"

namespace avs {

int stream_launch(int chain, bool is_v, int epi, int variant, const StreamParams& p, int sm_count, void* stream) {
    // scheduling variant (a per-plan option, AVIRB200_OPT_STREAM_VARIANT_H / _V); every variant
    // computes the same bits (variant 3, all rounds on the checked path, is host-emulation only)
    const int var = (variant >= 0 && variant < 3) ? variant : stream_default_variant(chain, is_v);
#define AVS_ROUTE(ID)                                                                              \
    case ID:                                                                                       \
        return is_v ? stream_launch_chain<ID, true>(var, epi, p, sm_count, stream)                 \
                    : stream_launch_chain<ID, false>(var, epi, p, sm_count, stream);
    switch (chain) {
        AVS_ROUTE(kChainDil24)
        AVS_ROUTE(kChainInl24)
        AVS_ROUTE(kChainInl3)
        AVS_ROUTE(kChainInl3D)
        AVS_ROUTE(kChainDil56)
        AVS_ROUTE(kChainUp2)
        AVS_ROUTE(kChainDil24Q)
    default: return -2;
    }
#undef AVS_ROUTE
}

} // namespace avs
