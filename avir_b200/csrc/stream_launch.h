// stream_launch.h -- engine-facing entry of the warp-streaming pass kernel (stream_pass.cu).
#pragma once

#include "stream_types.h"

namespace avs {

// Launches the chain kernel `chain` (StreamChainId) for one pass on `stream` (cudaStream_t).
// epi: column pass output stage, stream_epilogue_code() (ignored for the row pass).
// Returns 0 = launched, -2 = unknown chain, -1 = launch error.
int stream_launch(int chain, bool is_v, int epi, const StreamParams& p, void* stream);

// One chain's launcher, specialised in that chain's own translation unit (stream_chain.cu).
template <int ID>
int stream_launch_chain(bool is_v, int variant, int epi, const StreamParams& p, void* stream);
template <> int stream_launch_chain<kChainDil24>(bool, int, int, const StreamParams&, void*);
template <> int stream_launch_chain<kChainInl24>(bool, int, int, const StreamParams&, void*);
template <> int stream_launch_chain<kChainInl3>(bool, int, int, const StreamParams&, void*);
template <> int stream_launch_chain<kChainInl3D>(bool, int, int, const StreamParams&, void*);
template <> int stream_launch_chain<kChainDil56>(bool, int, int, const StreamParams&, void*);
template <> int stream_launch_chain<kChainUp2>(bool, int, int, const StreamParams&, void*);

} // namespace avs
