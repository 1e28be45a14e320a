// stream_launch.h -- engine-facing entry of the warp-streaming pass kernel (stream_pass.cu).
#pragma once

#include "stream_types.h"

namespace avs {

// Launches the chain kernel `chain` (StreamChainId) for one pass on `stream` (cudaStream_t).
// plain_f32: column pass whose destination is float without output gamma (store as is).
// Returns 0 = launched, -2 = unknown chain, -1 = launch error.
int stream_launch(int chain, bool is_v, bool plain_f32, const StreamParams& p, void* stream);

} // namespace avs
