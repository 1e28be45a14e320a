// stream_launch.h -- engine-facing entry of the warp-streaming pass kernel (stream_pass.cu).
#pragma once

#include "stream_types.h"

namespace avs {

// Launches the chain kernel `chain` (StreamChainId) for one pass on `stream` (cudaStream_t).
// epi: column pass output stage, stream_epilogue_code() (ignored for the row pass).
// Returns 0 = launched, -2 = unknown chain, -1 = launch error.
// variant: scheduling variant (stream_types.h), < 0 = the pass's default; sm_count: of the device
// the launch goes to (one persistent block per SM).
int stream_launch(int chain, bool is_v, int epi, int variant, const StreamParams& p, int sm_count, void* stream);

// One pass of one chain, specialised in its own translation unit (stream_chain.cu is compiled
// once per chain and pass: the kernels of a unit take minutes to compile, the units build in parallel).
template <int ID, bool IS_V>
int stream_launch_chain(int variant, int epi, const StreamParams& p, int sm_count, void* stream);
#define AVS_DECL_CHAIN(ID)                                                                  \
    template <> int stream_launch_chain<ID, false>(int, int, const StreamParams&, int, void*); \
    template <> int stream_launch_chain<ID, true>(int, int, const StreamParams&, int, void*);
AVS_DECL_CHAIN(kChainDil24)
AVS_DECL_CHAIN(kChainInl24)
AVS_DECL_CHAIN(kChainInl3)
AVS_DECL_CHAIN(kChainInl3D)
AVS_DECL_CHAIN(kChainDil56)
AVS_DECL_CHAIN(kChainUp2)
AVS_DECL_CHAIN(kChainDil24Q)
#undef AVS_DECL_CHAIN

} // namespace avs
