// fast_pass.cuh -- specialised pass kernel for 4-channel images (sm_100a).
//
// Same job as generic_pass.cuh (one kernel = one whole 1-D filtering chain for a tile of
// lines, every intermediate in shared memory) with the structure the FP32 pipe needs to be
// the limiter, because the bit-exact contract forbids FMA: each tap costs a separate
// multiply and add, so every other instruction in the inner loops competes for issue slots.
//
//   * lane = one channel PAIR (float2) of one line; a warp = the 32 lanes (16 lines) of ONE
//     position, so positions/taps/phases are warp-uniform: no divergence, tap reads are
//     shared-memory broadcasts, input reads are conflict-free 256-byte rows;
//   * register blocking: a thread produces M consecutive outputs from one register window
//     of inputs (window loads amortised over M x taps products) -- fully unrolled templates
//     for the chains of the BASELINE configs, plain loops for everything else;
//   * order-1 interpolation taps c0 + c1*x are row/column-invariant: they are formed once on
//     the host (same two float operations upstream performs) into an "effective phase"
//     table, so the kernels always run order-0 arithmetic;
//   * edge replication is materialised: a tile covers the UNCLAMPED index range its
//     consumer reads, out-of-domain positions hold the clamped sample, so inner loops
//     carry no index clamps;
//   * the per-tile tap rows are staged in shared memory once per step.
//
// Arithmetic order is upstream's (see generic_pass.cuh / oracle/avir_port.c); tests run
// every case through both kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdlib.h>

#include <map>
#include <utility>
#include <vector>

#include "device_plan.h"
#include "generic_pass.cuh"

namespace avb {

constexpr int kFastLanes = 32;      // lane pairs per position = 16 lines x 2
constexpr int kFastLines = 16;
constexpr int kFastThreads = 256;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastMaxSteps = 4;

enum FastVariant : int {
    kVarSimple = 0,
    kVarResizeDil24D2,   // float8_dil, FL 24, source step 2            (cfg3 mirror dil)
    kVarResizeDil56D4,   // float8_dil, FL 56, source step 4            (cfg5)
    kVarResizeDil32D2,
    kVarResizeInl18D2,   // interleaved, FL 18, source step 2           (cfg3 float4, cfg4)
    kVarResizeInl24D2,
    kVarFirDil8R1,       // float8_dil 8-tap (7 + pad) correction       (cfg3/cfg5 dil)
    kVarFirInl7R1,       // interleaved 7-tap (L = 3), R = 1            (LPF k=2, correction)
    kVarFirInl15R2,      // interleaved 15-tap (L = 7), R = 2           (cfg4 decimator)
};

struct FastStep {
    int kind, variant;
    int resample, latency, edge, ntaps, ntaps_pad;
    int out_len;
    int in_lo, in_hi;    // valid domain of the input line
    int upsampled, skip_odd, zero_start;
    const float* taps;   // FIR: ntaps floats; RESIZE: [n_eff][ntaps_pad]
    const int* src_pos;  // RESIZE
    const int* eff;      // RESIZE: per-output row of `taps`
};

struct FastAxis {
    int nsteps, src_len, dst_len;
    FastStep s[kFastMaxSteps];
};

struct FastParams {
    FastAxis ax;
    int is_v;
    int n_lines;          // rows (H) or pixel columns (V) in this launch
    int tile_out;
    int out0, out1;
    int span_a, span_b;   // shared rows of the two ping-pong buffers
    int taps_floats;      // shared floats reserved for staged taps
    const void* src;
    long long src_pitch;  // elements
    int src_type;
    int src_row_base;
    void* dst;
    long long dst_pitch;
    int dst_type;
    int dst_row_base;
    int gamma_in, gamma_out, alpha_index;
    float in_gamma_mult, out_gamma_mult;
    const float* srgb_lut;
    int round_mode;
    float tr_mul, tr_mul_inv, pk_out;
};

// ---- host+device range arithmetic (unclamped: tiles materialise edge replicas) ----------------

AVB_HD Range fast_input_range(const FastStep& s, Range o, const int* src_pos) {
    // o must lie inside the step's output domain
    Range r;
    if (s.kind == AVIRB200_STEP_FIR) {
        r.a = (o.a - s.edge) * s.resample - s.latency;
        r.b = (o.b - s.edge) * s.resample - s.latency + s.ntaps - 1;
    } else {
        const int d21 = s.ntaps / 2 - 1;
        r.a = src_pos[o.a] - d21;
        r.b = src_pos[o.b] - d21 + s.ntaps - 1;
        if (s.upsampled) {
            r.a >>= 1;
            r.b >>= 1;
        }
    }
    return r;
}

AVB_HD Range clampr(Range r, int lo, int hi) {
    Range c;
    c.a = imin(imax(r.a, lo), hi - 1);
    c.b = imin(imax(r.b, lo), hi - 1);
    return c;
}

// ---- small device helpers ------------------------------------------------------------------------

__device__ __forceinline__ float2 f2mul(float t, float2 x) {
    return make_float2(__fmul_rn(t, x.x), __fmul_rn(t, x.y));
}
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
__device__ __forceinline__ float2 f2hadd8(const float2* v) {
    return f2add(f2add(f2add(v[0], v[4]), f2add(v[1], v[5])),
                 f2add(f2add(v[2], v[6]), f2add(v[3], v[7])));
}

struct FTile {
    const float2* buf; // [pos - a][lane], pitch kFastPitch float2
    int a;
};

constexpr int kFastPitch = kFastLanes + 2; // float2 units: 272 bytes, keeps 16-byte alignment

__device__ __forceinline__ float2 ft(const FTile& t, int n, int lane) {
    return t.buf[(n - t.a) * kFastPitch + lane];
}

// Output stage for one element (gamma -> round -> clamp), 4-channel images.
__device__ __forceinline__ float epilogue_value_c4(const FastParams& p, float v, int c) {
    if (p.gamma_out) {
        if (c == p.alpha_index) v = __fmul_rn(v, p.out_gamma_mult);
        else v = __fmul_rn(lin2srgb(v), p.out_gamma_mult);
    }
    if (p.dst_type != AVIRB200_F32) {
        if (p.tr_mul == 1.0f) v = round_out(v, p.round_mode);
        else v = __fmul_rn(round_out(__fmul_rn(v, p.tr_mul_inv), p.round_mode), p.tr_mul);
        v = v < 0.0f ? 0.0f : (v > p.pk_out ? p.pk_out : v);
    }
    return v;
}

// De-interleaved RESIZE with long filters: group-major so that only 8 + (M-1)*D inputs and
// M x 8 lane accumulators are live at a time (a full register window would not fit).
template <int FL, int FLP, int D, int M>
__device__ __forceinline__ void resize_dil_groupmajor(const FTile& in, int p0, const float* tp,
                                                      int zero_start, int lane, float2* out) {
    constexpr int W = 8 + (M - 1) * D;
    float2 ln[M][8];
#pragma unroll
    for (int g = 0; g < FL / 8; ++g) {
        float2 x[W];
#pragma unroll
        for (int w = 0; w < W; ++w) x[w] = ft(in, p0 + g * 8 + w, lane);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const float* t = tp + m * FLP + g * 8;
            const float4 ta = *reinterpret_cast<const float4*>(t);
            const float4 tb = *reinterpret_cast<const float4*>(t + 4);
            const float tt[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float2 v = f2mul(tt[q], x[m * D + q]);
                ln[m][q] = (g == 0) ? v : f2add(ln[m][q], v);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float2 r = f2hadd8(ln[m]);
        if (zero_start) r = f2add(r, make_float2(0.0f, 0.0f));
        out[m] = r;
    }
}

// ---- blocked step routines: M outputs j .. j+M-1 from one register window -----------------------

// RESIZE, non-upsampled, uniform source step D.  `tp` = staged taps of output j (rows of
// ntaps_pad floats, consecutive outputs consecutive rows).
template <int SUM, int FL, int FLP, int D, int M>
__device__ __forceinline__ void resize_blocked(const FTile& in, int p0, const float* tp,
                                               int zero_start, int lane, float2* out) {
    constexpr int W = FL + (M - 1) * D;
    float2 x[W];
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = ft(in, p0 + w, lane);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const float* t = tp + m * FLP;
        float2 r;
        if (SUM == AVIRB200_SUM_DIL8) {
            float2 ln[8];
#pragma unroll
            for (int g = 0; g < FL / 8; ++g) {
                const float4 ta = *reinterpret_cast<const float4*>(t + g * 8);
                const float4 tb = *reinterpret_cast<const float4*>(t + g * 8 + 4);
                const float tt[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float2 v = f2mul(tt[q], x[m * D + g * 8 + q]);
                    ln[q] = (g == 0) ? v : f2add(ln[q], v);
                }
            }
            r = f2hadd8(ln);
        } else {
#pragma unroll
            for (int i = 0; i < FL; i += 2) {
                const float2 t2 = *reinterpret_cast<const float2*>(t + i);
                const float2 v0 = f2mul(t2.x, x[m * D + i]);
                r = (i == 0) ? v0 : f2add(r, v0);
                r = f2add(r, f2mul(t2.y, x[m * D + i + 1]));
            }
        }
        if (zero_start) r = f2add(r, make_float2(0.0f, 0.0f));
        out[m] = r;
    }
}

// FIR.  INL: folded symmetric form around the centre tap; DIL: full padded filter.
template <int SUM, int NT, int R, int M>
__device__ __forceinline__ void fir_blocked(const FTile& in, int p0, const float* taps, int lane,
                                            float2* out) {
    constexpr int W = NT + (M - 1) * R;
    float2 x[W];
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = ft(in, p0 + w, lane);
    float tt[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) tt[i] = taps[i];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        if (SUM == AVIRB200_SUM_DIL8) {
            float2 ln[8];
#pragma unroll
            for (int g = 0; g < NT / 8; ++g) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float2 v = f2mul(tt[g * 8 + q], x[m * R + g * 8 + q]);
                    ln[q] = (g == 0) ? v : f2add(ln[q], v);
                }
            }
            out[m] = f2hadd8(ln);
        } else {
            constexpr int L = NT / 2;
            float2 s = f2mul(tt[L], x[m * R + L]);
#pragma unroll
            for (int i = 1; i <= L; ++i)
                s = f2add(s, f2mul(tt[L + i], f2add(x[m * R + L + i], x[m * R + L - i])));
            out[m] = s;
        }
    }
}

// ---- plain-loop step routine: any geometry, one output ------------------------------------------

template <int SUM>
__device__ float2 step_simple(const FastStep& s, const FTile& in, int j, const float* tp, int lane) {
    if (s.kind == AVIRB200_STEP_FIR) {
        if (SUM == AVIRB200_SUM_INL) {
            const int L = s.latency;
            const int p = (j - s.edge) * s.resample;
            float2 sum = f2mul(tp[L], ft(in, p, lane));
            for (int i = 1; i <= L; ++i)
                sum = f2add(sum, f2mul(tp[L + i], f2add(ft(in, p + i, lane), ft(in, p - i, lane))));
            return sum;
        }
        const int p = (j - s.edge) * s.resample - s.latency;
        float2 ln[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) ln[q] = f2mul(tp[q], ft(in, p + q, lane));
        for (int i = 8; i < s.ntaps; i += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) ln[q] = f2add(ln[q], f2mul(tp[i + q], ft(in, p + i + q, lane)));
        }
        return f2hadd8(ln);
    }
    const int FL = s.ntaps;
    const int p = __ldg(s.src_pos + j) - (FL / 2 - 1);
    float2 r;
    if (SUM == AVIRB200_SUM_INL) {
        bool first = true;
        r = make_float2(0.0f, 0.0f);
        if (s.upsampled) {
            // only even virtual positions hold samples; upstream's doResize2 skips the rest
            for (int i = (p & 1); i < FL; i += 2) {
                const float2 v = f2mul(tp[i], ft(in, (p + i) >> 1, lane));
                r = first ? v : f2add(r, v);
                first = false;
            }
        } else {
            for (int i = 0; i < FL; ++i) {
                const float2 v = f2mul(tp[i], ft(in, p + i, lane));
                r = first ? v : f2add(r, v);
                first = false;
            }
        }
    } else {
        float2 ln[8];
        for (int i = 0; i < FL; i += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float2 v = f2mul(tp[i + q], ft(in, p + i + q, lane));
                ln[q] = (i == 0) ? v : f2add(ln[q], v);
            }
        }
        r = f2hadd8(ln);
    }
    if (s.zero_start) r = f2add(r, make_float2(0.0f, 0.0f));
    return r;
}

// ---- the kernel ------------------------------------------------------------------------------------

template <int SUM>
__global__ void __launch_bounds__(kFastThreads, 2)
fast_pass_kernel(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + (size_t)p.span_a * kFastPitch;
    float* stap = reinterpret_cast<float*>(bufB + (size_t)p.span_b * kFastPitch);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int line0 = blockIdx.y * kFastLines;
    const int nlines = imin(kFastLines, p.n_lines - line0);
    const int j0 = p.out0 + blockIdx.x * p.tile_out;
    const int j1 = imin(j0 + p.tile_out, p.out1) - 1;
    const int ns = p.ax.nsteps;

    // ranges: rng[i] = unclamped positions of step i's INPUT held in shared memory
    Range rng[kFastMaxSteps + 1];
    rng[ns].a = j0;
    rng[ns].b = j1;
    for (int i = ns - 1; i >= 0; --i) {
        const FastStep& s = p.ax.s[i];
        rng[i] = fast_input_range(s, clampr(rng[i + 1], 0, s.out_len), s.src_pos);
    }

    // ---- stage the source tile (edge replicas materialised)
    {
        const int a = rng[0].a, n = rng[0].b - rng[0].a + 1;
        if (p.is_v) {
            // a "line" is a pixel column; 16 pixels = 64 contiguous floats of a row
            const float* src = static_cast<const float*>(p.src);
            const int q = tid & 15;  // pixel within the strip (float4)
            const int r0 = tid >> 4; // 16 rows per sweep
            const bool ok = q < nlines;
            for (int pos = r0; pos < n; pos += kFastThreads / 16) {
                int y = imin(imax(a + pos, 0), p.ax.src_len - 1) - p.src_row_base;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)y * p.src_pitch) + line0 + q);
                *reinterpret_cast<float4*>(bufA + pos * kFastPitch + q * 2) = v;
            }
        } else {
            const int px = tid & 31; // 32 consecutive positions per sweep
            const int r0 = tid >> 5; // 8 rows per sweep
            for (int r = r0; r < kFastLines; r += kFastThreads / 32) {
                const bool ok = r < nlines;
                const size_t rowoff = (size_t)(line0 + (ok ? r : 0)) * p.src_pitch;
                for (int pos = px; pos < n; pos += 32) {
                    const int x = imin(imax(a + pos, 0), p.ax.src_len - 1);
                    float4 v;
                    if (p.src_type == AVIRB200_F32) {
                        v = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.src) + rowoff) + x);
                    } else if (p.src_type == AVIRB200_U8) {
                        const uchar4 b = __ldg(reinterpret_cast<const uchar4*>(static_cast<const unsigned char*>(p.src) + rowoff) + x);
                        v = make_float4((float)b.x, (float)b.y, (float)b.z, (float)b.w);
                        if (p.gamma_in) {
                            const int ai = p.alpha_index;
                            v.x = (ai == 0) ? __fmul_rn(v.x, p.in_gamma_mult) : p.srgb_lut[b.x];
                            v.y = p.srgb_lut[b.y];
                            v.z = p.srgb_lut[b.z];
                            v.w = (ai == 3) ? __fmul_rn(v.w, p.in_gamma_mult) : p.srgb_lut[b.w];
                        }
                    } else {
                        const ushort4 b = __ldg(reinterpret_cast<const ushort4*>(static_cast<const unsigned short*>(p.src) + rowoff) + x);
                        v = make_float4((float)b.x, (float)b.y, (float)b.z, (float)b.w);
                    }
                    if (p.gamma_in && p.src_type != AVIRB200_U8) {
                        const int ai = p.alpha_index;
                        v.x = (ai == 0) ? __fmul_rn(v.x, p.in_gamma_mult) : srgb2lin(v.x, p.in_gamma_mult);
                        v.y = srgb2lin(v.y, p.in_gamma_mult);
                        v.z = srgb2lin(v.z, p.in_gamma_mult);
                        v.w = (ai == 3) ? __fmul_rn(v.w, p.in_gamma_mult) : srgb2lin(v.w, p.in_gamma_mult);
                    }
                    *reinterpret_cast<float4*>(bufA + pos * kFastPitch + r * 2) = v;
                }
            }
        }
    }

    // ---- the chain
    for (int i = 0; i < ns; ++i) {
        const FastStep& s = p.ax.s[i];
        FTile in;
        in.buf = (i & 1) ? bufB : bufA;
        in.a = rng[i].a;
        float2* ob = (i & 1) ? bufA : bufB;
        const int oa = rng[i + 1].a, on = rng[i + 1].b - rng[i + 1].a + 1;
        const bool last = (i == ns - 1);
        const Range dom = clampr(rng[i + 1], 0, s.out_len); // outputs actually computed

        // stage taps: FIR -> ntaps floats; RESIZE -> one row per in-domain output
        if (s.kind == AVIRB200_STEP_FIR) {
            for (int t = tid; t < s.ntaps; t += kFastThreads) stap[t] = __ldg(s.taps + t);
        } else {
            const int rows = dom.b - dom.a + 1;
            const int fl4 = s.ntaps_pad / 4;
            for (int t = tid; t < rows * fl4; t += kFastThreads) {
                const int rr = t / fl4, c4 = t - rr * fl4;
                const int e = __ldg(s.eff + dom.a + rr);
                reinterpret_cast<float4*>(stap)[t] =
                    __ldg(reinterpret_cast<const float4*>(s.taps + (size_t)e * s.ntaps_pad) + c4);
            }
        }
        __syncthreads(); // source tile / previous step output and taps are in place

        // warp w handles a contiguous chunk of positions
        int chunk = (on + kFastWarps - 1) / kFastWarps;
        chunk = (chunk + 3) & ~3;
        const int jb = oa + warp * chunk;
        const int je = imin(jb + chunk, oa + on);

        auto emit = [&](int j, float2 v) {
            if (last && p.is_v) {
                const int c0 = (lane & 1) * 2;
                v.x = epilogue_value_c4(p, v.x, c0);
                v.y = epilogue_value_c4(p, v.y, c0 + 1);
                const int px = lane >> 1;
                if (px < nlines) {
                    const size_t g = (size_t)(j - p.dst_row_base) * p.dst_pitch +
                                     (size_t)(line0 + px) * 4 + c0;
                    if (p.dst_type == AVIRB200_F32) {
                        *reinterpret_cast<float2*>(static_cast<float*>(p.dst) + g) = v;
                    } else if (p.dst_type == AVIRB200_U8) {
                        *reinterpret_cast<uchar2*>(static_cast<unsigned char*>(p.dst) + g) =
                            make_uchar2((unsigned char)v.x, (unsigned char)v.y);
                    } else {
                        *reinterpret_cast<ushort2*>(static_cast<unsigned short*>(p.dst) + g) =
                            make_ushort2((unsigned short)v.x, (unsigned short)v.y);
                    }
                }
            } else {
                ob[(j - oa) * kFastPitch + lane] = v;
            }
        };

        int j = jb;
        while (j < je) {
            // blocked fast path: M in-domain outputs with the geometry the variant expects
            bool done = false;
            if (s.variant != kVarSimple && j >= dom.a && j + 3 <= dom.b && j + 3 < je) {
                float2 o4[4];
                if (s.kind == AVIRB200_STEP_RESIZE) {
                    const int sp0 = __ldg(s.src_pos + j);
                    const int sp3 = __ldg(s.src_pos + j + 3);
                    const int sp1 = __ldg(s.src_pos + j + 1);
                    const float* tp = stap + (size_t)(j - dom.a) * s.ntaps_pad;
                    const int p0 = sp0 - (s.ntaps / 2 - 1);
                    if (s.variant == kVarResizeDil24D2 && sp3 - sp0 == 6 && sp1 - sp0 == 2) {
                        resize_blocked<AVIRB200_SUM_DIL8, 24, 24, 2, 4>(in, p0, tp, s.zero_start, lane, o4);
                        done = true;
                    } else if (s.variant == kVarResizeDil32D2 && sp3 - sp0 == 6 && sp1 - sp0 == 2) {
                        resize_blocked<AVIRB200_SUM_DIL8, 32, 32, 2, 4>(in, p0, tp, s.zero_start, lane, o4);
                        done = true;
                    } else if (s.variant == kVarResizeInl18D2 && sp3 - sp0 == 6 && sp1 - sp0 == 2) {
                        resize_blocked<AVIRB200_SUM_INL, 18, 20, 2, 4>(in, p0, tp, s.zero_start, lane, o4);
                        done = true;
                    } else if (s.variant == kVarResizeInl24D2 && sp3 - sp0 == 6 && sp1 - sp0 == 2) {
                        resize_blocked<AVIRB200_SUM_INL, 24, 24, 2, 4>(in, p0, tp, s.zero_start, lane, o4);
                        done = true;
                    } else if (s.variant == kVarResizeDil56D4 && sp3 - sp0 == 12 && sp1 - sp0 == 4) {
                        resize_dil_groupmajor<56, 56, 4, 2>(in, p0, tp, s.zero_start, lane, o4);
                        resize_dil_groupmajor<56, 56, 4, 2>(in, p0 + 8, tp + 2 * 56, s.zero_start, lane, o4 + 2);
                        done = true;
                    }
                } else {
                    if (s.variant == kVarFirDil8R1) {
                        fir_blocked<AVIRB200_SUM_DIL8, 8, 1, 4>(in, (j - s.edge) - s.latency, stap, lane, o4);
                        done = true;
                    } else if (s.variant == kVarFirInl7R1) {
                        fir_blocked<AVIRB200_SUM_INL, 7, 1, 4>(in, (j - s.edge) - 3, stap, lane, o4);
                        done = true;
                    } else if (s.variant == kVarFirInl15R2) {
                        fir_blocked<AVIRB200_SUM_INL, 15, 2, 4>(in, (j - s.edge) * 2 - 7, stap, lane, o4);
                        done = true;
                    }
                }
                if (done) {
#pragma unroll
                    for (int m = 0; m < 4; ++m) emit(j + m, o4[m]);
                    j += 4;
                }
            }
            if (!done) {
                const int jj = imin(imax(j, dom.a), dom.b); // edge replica: value of the clamped output
                const float* tp = (s.kind == AVIRB200_STEP_FIR) ? stap
                                                                : stap + (size_t)(jj - dom.a) * s.ntaps_pad;
                emit(j, step_simple<SUM>(s, in, jj, tp, lane));
                ++j;
            }
        }
        __syncthreads();
    }

    if (!p.is_v) {
        // coalesced store of the row-pass tile: [pos][row] in shared -> rows of float4 pixels
        const float2* ob = (ns & 1) ? bufB : bufA;
        const int oa = rng[ns].a, on = rng[ns].b - rng[ns].a + 1;
        const int px = tid & 31, r0 = tid >> 5;
        for (int r = r0; r < nlines; r += kFastThreads / 32) {
            float4* drow = reinterpret_cast<float4*>(static_cast<float*>(p.dst) +
                                                     (size_t)(line0 + r) * p.dst_pitch);
            for (int pos = px; pos < on; pos += 32)
                drow[oa + pos] = *reinterpret_cast<const float4*>(ob + pos * kFastPitch + r * 2);
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------

struct FastPass {
    bool ok = false;
    FastAxis ax;        // device pointers
    FastAxis hax;       // host pointers (range arithmetic)
    int tile_out = 0, span_a = 0, span_b = 0, taps_floats = 0;
    size_t smem = 0;
    std::vector<std::vector<float> > eff_taps;
    std::vector<std::vector<int> > eff_idx, src_pos;
    void* arena = nullptr;
};

struct FastPlan {
    bool h_ok = false, v_ok = false;
    FastPass h, v;
};

inline bool env_fast_enabled() {
    static const bool on = [] {
        const char* e = getenv("AVIRB200_DISABLE_FAST");
        return !(e && e[0] == '1');
    }();
    return on;
}

const size_t kFastSmemBudget = 100 * 1024; // two resident blocks per SM

// Largest shared-memory footprint over all tiles for a given tile size.
inline bool fast_footprint(const FastAxis& hax, int t, int out0, int out1, int* span_a, int* span_b,
                           int* taps_floats) {
    int sa = 0, sb = 0, tf = 0;
    const int ns = hax.nsteps;
    for (int j0 = out0; j0 < out1; j0 += t) {
        Range r{j0, imin(j0 + t, out1) - 1};
        for (int i = ns - 1; i >= 0; --i) {
            const FastStep& s = hax.s[i];
            // output tile of step i lives in buffer ((i+1)&1 ? B : A)
            const int on = r.b - r.a + 1;
            if ((i + 1) & 1) sb = imax(sb, on); else sa = imax(sa, on);
            const Range dom = clampr(r, 0, s.out_len);
            const int need = (s.kind == AVIRB200_STEP_FIR) ? s.ntaps
                                                           : (dom.b - dom.a + 1) * s.ntaps_pad;
            tf = imax(tf, need);
            r = fast_input_range(s, dom, s.src_pos);
        }
        sa = imax(sa, r.b - r.a + 1); // source tile in A
    }
    *span_a = sa; *span_b = sb; *taps_floats = (tf + 3) & ~3;
    return true;
}

inline size_t fast_smem_bytes(int sa, int sb, int tf) {
    return ((size_t)sa + sb) * kFastPitch * sizeof(float2) + (size_t)tf * sizeof(float);
}

inline void fast_choose_tile(FastPass& fp, int out0, int out1) {
    static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16, 8, 4};
    for (int t : cand) {
        int sa, sb, tf;
        fast_footprint(fp.hax, t, out0, out1, &sa, &sb, &tf);
        if (fast_smem_bytes(sa, sb, tf) <= kFastSmemBudget || t == 4) {
            fp.tile_out = t; fp.span_a = sa; fp.span_b = sb; fp.taps_floats = tf;
            fp.smem = fast_smem_bytes(sa, sb, tf);
            return;
        }
    }
}

// Builds the fast description of one axis from the (host-pointer) generic one.  Returns
// false when the chain is outside the fast kernel's scope (filtered upsample, zero-stuffed
// de-interleaved resize, too many steps); the generic kernel then runs it.
inline bool fast_build_axis(FastPass& fp, const DevAxis& host, int sum_mode) {
    if (host.nsteps > kFastMaxSteps) return false;
    FastAxis& a = fp.hax;
    a.nsteps = host.nsteps; a.src_len = host.src_len; a.dst_len = host.dst_len;
    fp.eff_taps.assign(host.nsteps, {});
    fp.eff_idx.assign(host.nsteps, {});
    fp.src_pos.assign(host.nsteps, {});
    for (int i = 0; i < host.nsteps; ++i) {
        const DevStep& d = host.steps[i];
        FastStep& s = a.s[i];
        s.kind = d.kind; s.variant = kVarSimple;
        s.resample = d.resample; s.latency = d.latency; s.edge = d.edge;
        s.ntaps = d.ntaps; s.ntaps_pad = (d.ntaps + 3) & ~3;
        s.out_len = d.out_len; s.in_lo = d.in_lo; s.in_hi = d.in_hi;
        s.upsampled = d.upsampled; s.skip_odd = d.skip_odd; s.zero_start = d.zero_start;
        s.taps = nullptr; s.src_pos = nullptr; s.eff = nullptr;
        if (d.kind == AVIRB200_STEP_UPSAMPLE) return false;
        if (d.in_lo != 0) return false;
        if (d.kind == AVIRB200_STEP_FIR) {
            if (sum_mode == AVIRB200_SUM_INL && d.ntaps != 2 * d.latency + 1) return false;
            if (sum_mode == AVIRB200_SUM_DIL8 && (d.ntaps & 7)) return false;
            fp.eff_taps[i].assign(d.taps, d.taps + d.ntaps);
            if (sum_mode == AVIRB200_SUM_DIL8 && d.ntaps == 8 && d.resample == 1) s.variant = kVarFirDil8R1;
            if (sum_mode == AVIRB200_SUM_INL && d.ntaps == 7 && d.resample == 1) s.variant = kVarFirInl7R1;
            if (sum_mode == AVIRB200_SUM_INL && d.ntaps == 15 && d.resample == 2) s.variant = kVarFirInl15R2;
        } else {
            if (d.upsampled && !(sum_mode == AVIRB200_SUM_INL && d.skip_odd)) return false;
            if (sum_mode == AVIRB200_SUM_DIL8 && (d.ntaps & 7)) return false;
            // effective phases: (phase, frac) -> c0 + c1*frac, the two float operations
            // upstream performs per tap (avir.h:3945, avir_dil.h:649-650)
            std::map<std::pair<int, uint32_t>, int> seen;
            fp.eff_idx[i].resize(d.out_len);
            fp.src_pos[i].assign(d.src_pos, d.src_pos + d.out_len);
            const int FL = d.ntaps, FLP = s.ntaps_pad;
            for (int j = 0; j < d.out_len; ++j) {
                uint32_t fb = 0;
                if (d.order) memcpy(&fb, &d.frac[j], 4);
                const std::pair<int, uint32_t> key(d.phase[j], fb);
                auto it = seen.find(key);
                if (it == seen.end()) {
                    const int row = (int)seen.size();
                    it = seen.emplace(key, row).first;
                    const float* c0 = d.taps + (size_t)d.phase[j] * FL * (d.order + 1);
                    const float x = d.frac[j];
                    fp.eff_taps[i].resize((size_t)(row + 1) * FLP, 0.0f);
                    float* o = &fp.eff_taps[i][(size_t)row * FLP];
                    for (int t = 0; t < FL; ++t) {
                        if (d.order) {
                            volatile float prod = c0[FL + t] * x; // keep the two roundings apart
                            o[t] = c0[t] + prod;
                        } else {
                            o[t] = c0[t];
                        }
                    }
                }
                fp.eff_idx[i][j] = it->second;
            }
            if (!d.upsampled) {
                if (sum_mode == AVIRB200_SUM_DIL8 && FL == 24) s.variant = kVarResizeDil24D2;
                if (sum_mode == AVIRB200_SUM_DIL8 && FL == 32) s.variant = kVarResizeDil32D2;
                if (sum_mode == AVIRB200_SUM_DIL8 && FL == 56) s.variant = kVarResizeDil56D4;
                if (sum_mode == AVIRB200_SUM_INL && FL == 18) s.variant = kVarResizeInl18D2;
                if (sum_mode == AVIRB200_SUM_INL && FL == 24) s.variant = kVarResizeInl24D2;
            }
        }
    }
    return true;
}

inline size_t fa_align(size_t v) { return (v + 255) / 256 * 256; }

inline int fast_upload(FastPass& fp) {
    size_t bytes = 0;
    const int ns = fp.hax.nsteps;
    for (int i = 0; i < ns; ++i)
        bytes += fa_align(fp.eff_taps[i].size() * 4) + fa_align(fp.eff_idx[i].size() * 4) +
                 fa_align(fp.src_pos[i].size() * 4);
    if (cudaMalloc(&fp.arena, bytes + 256) != cudaSuccess) return -1;
    std::vector<char> img(bytes + 256, 0);
    size_t off = 0;
    fp.ax = fp.hax;
    char* base = static_cast<char*>(fp.arena);
    for (int i = 0; i < ns; ++i) {
        auto put = [&](const void* src, size_t n) -> const void* {
            if (n == 0) return nullptr;
            memcpy(img.data() + off, src, n);
            const void* d = base + off;
            off += fa_align(n);
            return d;
        };
        fp.ax.s[i].taps = static_cast<const float*>(put(fp.eff_taps[i].data(), fp.eff_taps[i].size() * 4));
        fp.ax.s[i].eff = static_cast<const int*>(put(fp.eff_idx[i].data(), fp.eff_idx[i].size() * 4));
        fp.ax.s[i].src_pos = static_cast<const int*>(put(fp.src_pos[i].data(), fp.src_pos[i].size() * 4));
        fp.hax.s[i].taps = fp.eff_taps[i].data();
        fp.hax.s[i].eff = fp.eff_idx[i].empty() ? nullptr : fp.eff_idx[i].data();
        fp.hax.s[i].src_pos = fp.src_pos[i].empty() ? nullptr : fp.src_pos[i].data();
    }
    if (cudaMemcpy(fp.arena, img.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    return 0;
}

inline void fast_plan_init(FastPlan& f, const DevAxis& h_host, const DevAxis& v_host,
                           const avirb200_plan_desc& d) {
    if (d.channels != 4) return;
    FastPass* ps[2] = {&f.h, &f.v};
    const DevAxis* hs[2] = {&h_host, &v_host};
    for (int a = 0; a < 2; ++a) {
        FastPass& fp = *ps[a];
        if (!fast_build_axis(fp, *hs[a], d.sum_mode)) continue;
        // host pointers for range arithmetic first, then upload
        for (int i = 0; i < fp.hax.nsteps; ++i)
            fp.hax.s[i].src_pos = fp.src_pos[i].empty() ? nullptr : fp.src_pos[i].data();
        if (fast_upload(fp) != 0) continue;
        fast_choose_tile(fp, 0, hs[a]->dst_len);
        fp.ok = true;
    }
    f.h_ok = f.h.ok;
    f.v_ok = f.v.ok;
}

inline void fast_plan_free(FastPlan& f) {
    cudaFree(f.h.arena);
    cudaFree(f.v.arena);
    f.h.arena = f.v.arena = nullptr;
}

inline void fast_fill_common(FastParams& p, const avirb200_plan_desc& d, const float* lut) {
    p.gamma_in = (d.use_gamma & 1) ? 1 : 0;
    p.gamma_out = (d.use_gamma & 2) ? 1 : 0;
    p.alpha_index = d.alpha_index;
    p.in_gamma_mult = d.in_gamma_mult;
    p.out_gamma_mult = d.out_gamma_mult;
    p.srgb_lut = lut;
    p.round_mode = d.round_mode;
    p.tr_mul = d.tr_mul;
    p.tr_mul_inv = d.tr_mul_inv;
    p.pk_out = d.pk_out;
}

inline int fast_launch(const FastParams& p, size_t smem, int sum_mode, cudaStream_t st) {
    dim3 grid((p.out1 - p.out0 + p.tile_out - 1) / p.tile_out, (p.n_lines + kFastLines - 1) / kFastLines);
    if (grid.y > 65535) return -1;
    cudaError_t e;
    if (sum_mode == AVIRB200_SUM_DIL8) {
        e = cudaFuncSetAttribute(fast_pass_kernel<AVIRB200_SUM_DIL8>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -1;
        fast_pass_kernel<AVIRB200_SUM_DIL8><<<grid, kFastThreads, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(fast_pass_kernel<AVIRB200_SUM_INL>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -1;
        fast_pass_kernel<AVIRB200_SUM_INL><<<grid, kFastThreads, smem, st>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

inline size_t fast_elsize(int t) { return t == AVIRB200_U8 ? 1 : (t == AVIRB200_U16 ? 2 : 4); }

// Returns 0 = launched, -2 = not applicable (alignment: the caller runs the generic kernel),
// -1 = launch error.
inline int fast_row_pass(const FastPlan& f, const avirb200_plan_desc& d, const void* d_src,
                         size_t src_pitch, float* d_mid, int rows, const float* lut, cudaStream_t st) {
    const size_t es = fast_elsize(d.in_type);
    if (((uintptr_t)d_src % (4 * es)) != 0 || (src_pitch % 4) != 0 || ((uintptr_t)d_mid % 16) != 0)
        return -2;
    FastParams p;
    memset(&p, 0, sizeof p);
    fast_fill_common(p, d, lut);
    p.ax = f.h.ax;
    p.is_v = 0;
    p.n_lines = rows;
    p.tile_out = f.h.tile_out;
    p.out0 = 0; p.out1 = d.dst_w;
    p.span_a = f.h.span_a; p.span_b = f.h.span_b; p.taps_floats = f.h.taps_floats;
    p.src = d_src; p.src_pitch = (long long)src_pitch; p.src_type = d.in_type;
    p.dst = d_mid; p.dst_pitch = (long long)d.dst_w * 4; p.dst_type = AVIRB200_F32;
    return fast_launch(p, f.h.smem, d.sum_mode, st);
}

inline int fast_col_pass(const FastPlan& f, const avirb200_plan_desc& d, const float* d_mid,
                         int mid_row_base, void* d_dst, size_t dst_pitch, int out0, int out1,
                         const float* lut, cudaStream_t st) {
    const size_t es = fast_elsize(d.out_type);
    if (((uintptr_t)d_dst % (2 * es)) != 0 || (dst_pitch % 2) != 0 || ((uintptr_t)d_mid % 16) != 0)
        return -2;
    FastParams p;
    memset(&p, 0, sizeof p);
    fast_fill_common(p, d, lut);
    p.ax = f.v.ax;
    p.is_v = 1;
    p.n_lines = d.dst_w;
    int tile = f.v.tile_out, sa = f.v.span_a, sb = f.v.span_b, tf = f.v.taps_floats;
    size_t smem = f.v.smem;
    if (out0 != 0 || out1 != d.dst_h) { // a shard: footprint of its own tiles
        fast_footprint(f.v.hax, tile, out0, out1, &sa, &sb, &tf);
        smem = fast_smem_bytes(sa, sb, tf);
        if (smem > 200 * 1024) return -2;
    }
    p.tile_out = tile;
    p.out0 = out0; p.out1 = out1;
    p.span_a = sa; p.span_b = sb; p.taps_floats = tf;
    p.src = d_mid; p.src_pitch = (long long)d.dst_w * 4; p.src_type = AVIRB200_F32;
    p.src_row_base = mid_row_base;
    p.dst = d_dst; p.dst_pitch = (long long)dst_pitch; p.dst_type = d.out_type;
    p.dst_row_base = out0;
    return fast_launch(p, smem, d.sum_mode, st);
}

} // namespace avb
