// device_plan.h -- device-side mirror of the C-ABI plan descriptor (avirb200.h) and the
// range arithmetic shared by host launch code and kernels.
#pragma once

#include <stdint.h>

#include "avirb200.h"

#if defined(__CUDACC__)
#define AVB_HD __host__ __device__ __forceinline__
#else
#define AVB_HD inline
#endif

namespace avb {

struct DevStep {
    int kind, resample, latency, edge;
    int in_len, out_len, ntaps, order;
    int upsampled, skip_odd, zero_start, nphases;
    int out_prefix, out_suffix, in_prefix, in_suffix;
    int n_prefix_dc, n_suffix_dc;
    int in_lo, in_hi; // valid index range [in_lo, in_hi) of this step's INPUT line
    const float* taps;
    const int* src_pos;
    const int* phase;
    const float* frac;
    const float* prefix_dc;
    const float* suffix_dc;
};

struct DevAxis {
    int src_len, dst_len, nsteps;
    DevStep steps[AVIRB200_MAX_STEPS];
};

struct Range {
    int a, b; // inclusive
};

AVB_HD int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
AVB_HD int imin(int a, int b) { return a < b ? a : b; }
AVB_HD int imax(int a, int b) { return a > b ? a : b; }

// Input indices (of the step's own input line) needed to produce outputs [o.a, o.b].
// src_pos must be readable from the caller's address space.
AVB_HD Range step_input_range(const DevStep& s, Range o, const int* src_pos) {
    Range r;
    if (s.kind == AVIRB200_STEP_FIR) {
        r.a = (o.a - s.edge) * s.resample - s.latency;
        r.b = (o.b - s.edge) * s.resample - s.latency + s.ntaps - 1;
    } else if (s.kind == AVIRB200_STEP_RESIZE) {
        const int d21 = s.ntaps / 2 - 1;
        r.a = src_pos[o.a] - d21;
        r.b = src_pos[o.b] - d21 + s.ntaps - 1;
        if (s.upsampled) {
            r.a >>= 1;
            r.b >>= 1;
        }
    } else {
        const int R = s.resample;
        r.a = floordiv(o.a + s.latency - (s.ntaps - 1) + R - 1, R);
        r.b = floordiv(o.b + s.latency, R);
        const int pfx = -s.in_prefix * R;
        const int sfx = (s.in_len + s.in_suffix) * R - s.latency;
        if (o.a < pfx + s.n_prefix_dc && o.b >= pfx) r.a = imin(r.a, 0);
        if (o.a < sfx + s.n_suffix_dc && o.b >= sfx) r.b = imax(r.b, s.in_len - 1);
    }
    r.a = imin(imax(r.a, s.in_lo), s.in_hi - 1);
    r.b = imin(imax(r.b, s.in_lo), s.in_hi - 1);
    return r;
}

// Valid index range of a step's OUTPUT line.
AVB_HD Range step_output_domain(const DevStep& s) {
    Range r;
    if (s.kind == AVIRB200_STEP_UPSAMPLE) {
        r.a = -s.out_prefix;
        r.b = s.out_len + s.out_suffix - 1;
    } else {
        r.a = 0;
        r.b = s.out_len - 1;
    }
    return r;
}

} // namespace avb
