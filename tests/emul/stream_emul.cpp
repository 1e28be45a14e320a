// stream_emul.cpp -- TEST INFRASTRUCTURE: lockstep host emulation of the warp-streaming pass
// kernel (avir_b200/csrc/stream_kernel.cuh).
//
// The kernel source is compiled here for the host: a warp is 32 threads that meet at a
// pthread barrier wherever the device code executes __syncwarp(), cp.async becomes an
// immediate 16-byte copy, "shared memory" is a per-warp heap block.  The emulation exists so
// that the kernel's index logic (rings, software-pipeline delays, run splitting, edge
// batches, shard ranges) can be checked bit-for-bit against the oracle on a machine without
// a GPU (tests/test_stream_emul.py).  It is never linked into libavirb200.so and is not a
// CPU execution path of the product.
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC (tests/emul/build.py).

#include <pthread.h>
#include <stdlib.h>

#include <thread>
#include <vector>

#include <atomic>

static thread_local pthread_barrier_t* g_bar = nullptr;
void avs_emul_syncwarp() { pthread_barrier_wait(g_bar); }
static std::atomic<long> g_oob(0);
namespace avs {
thread_local const unsigned char* avs_emul_src_lo = nullptr;
thread_local const unsigned char* avs_emul_src_hi = nullptr;
thread_local const unsigned char* avs_emul_alt_lo[2] = {nullptr, nullptr};
thread_local const unsigned char* avs_emul_alt_hi[2] = {nullptr, nullptr};
void avs_emul_count_oob() { ++g_oob; }
unsigned long long avs_emul_xs_add(unsigned long long* c, unsigned long long v) { // (one lane of one warp at a time)
    const unsigned long long old = *c;
    *c = old + v;
    return old;
}
} // namespace avs
static const unsigned char* g_lo = nullptr; // bounds of the pass's source buffer (copied into every lane thread)
static const unsigned char* g_hi = nullptr;
static const unsigned char* g_alt_lo[2] = {nullptr, nullptr};
static const unsigned char* g_alt_hi[2] = {nullptr, nullptr};

#include "stream_kernel.cuh"

using namespace avs;

namespace {

template <class C, bool IS_V, int EPI>
void emul_pass(const StreamParams& p, int nwarps) {
    const size_t sm_float2 = (size_t)(IS_V ? C::WARP_F2_V : C::WARP_F2_H);
    for (int gw = 0; gw < nwarps; ++gw) {
        std::vector<float2> sm(sm_float2);
        // poison: reads of never-written slots must not go unnoticed
        for (auto& v : sm) v = make_float2(__builtin_nanf(""), __builtin_nanf(""));
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, 32);
        std::vector<std::thread> th;
        for (int lane = 0; lane < 32; ++lane) {
            th.emplace_back([&, lane] {
                g_bar = &bar;
                avs_emul_src_lo = g_lo;
                avs_emul_src_hi = g_hi;
                for (int q = 0; q < 2; ++q) { avs_emul_alt_lo[q] = g_alt_lo[q]; avs_emul_alt_hi[q] = g_alt_hi[q]; }
                stream_warp_main<C, IS_V, EPI>(p, gw, nwarps, lane, sm.data(), p.srgb_lut);
            });
        }
        for (auto& t : th) t.join();
        pthread_barrier_destroy(&bar);
    }
}

template <bool IS_V>
bool emul_dispatch(int chain, int variant, const StreamParams& p, int nwarps, int epi) {
    return stream_dispatch(chain, IS_V, variant, p.src_type, [&](auto tag, auto pass) {
        using C = typename decltype(tag)::type;
        if constexpr (decltype(pass)::is_v != IS_V) {
            (void)p; // (the dispatcher instantiates the callback for both passes)
        } else if constexpr (IS_V) {
            if (epi == 1) emul_pass<C, true, 1>(p, nwarps);
            else if (epi == 2) emul_pass<C, true, 2>(p, nwarps);
            else emul_pass<C, true, 0>(p, nwarps);
        } else {
            if (p.xs_count != nullptr) emul_pass<C, false, kEpiXs>(p, nwarps);
            else emul_pass<C, false, 0>(p, nwarps);
        }
    });
}

} // namespace

extern "C" {

// 1 when both axes of the descriptor run on the streaming kernel (f32 source only).
int stream_emul_applicable(const avirb200_plan_desc* d) {
    StreamAxisPlan h, v;
    return stream_row_source_ok(*d) && stream_plan_axis(d->h, d->sum_mode, d->channels, h, 1) &&
           stream_plan_axis(d->v, d->sum_mode, d->channels, v, 1);
}

// Row pass with `warps_h` emulated warps, then the column pass in `bands` destination bands
// (as the sharded schedule runs it) with `warps_v` warps each.
// lut: the 256-entry u8 sRGB linearisation table (read for sRGB sources only).
// allow: 1 = every chain (as AVIRB200_OPT_ALL_STREAM_CHAINS = 1), 2 = the 4-output-batch twin of the
// headline chain where it applies.
// need: per band (need_row0, need_rows) of avirb200_shard_query_desc, or null.  With it every band's
// column pass gets a buffer holding exactly those intermediate rows, as a rank of the sharded
// schedule does, and every source read is bounds-checked: returns -5 if any read left its buffer.
int stream_emul_resize(const avirb200_plan_desc* d, const void* src, size_t src_pitch, void* dst,
                       size_t dst_pitch, int warps_h, int warps_v, int bands, int variant, const float* lut, int allow,
                       const int* need, int seg_top, int seg_bot) {
    StreamAxisPlan h, v;
    if (!stream_row_source_ok(*d) || !stream_plan_axis(d->h, d->sum_mode, d->channels, h, allow) ||
        !stream_plan_axis(d->v, d->sum_mode, d->channels, v, allow))
        return -4;
    std::vector<float> mid((size_t)d->src_h * d->dst_w * 4);
    StreamParams p;
    stream_fill_params(p, h, *d);
    p.n_lines = d->src_h;
    p.out0 = 0;
    p.out1 = d->dst_w;
    p.src = src;
    p.src_type = stream_row_source_code(*d);
    p.srgb_lut = lut;
    p.src_pitch = (long long)src_pitch;
    p.dst = mid.data();
    p.dst_pitch = (long long)d->dst_w * 4;
    p.dst_type = AVIRB200_F32;
    g_oob = 0;
    {
        const size_t esz = (d->in_type == AVIRB200_U8) ? 1 : (d->in_type == AVIRB200_U16 ? 2 : 4);
        g_lo = static_cast<const unsigned char*>(src);
        g_hi = g_lo + ((size_t)(d->src_h - 1) * src_pitch + (size_t)d->src_w * 4) * esz;
    }
    if (seg_top + seg_bot > 0 && seg_top + seg_bot < d->src_h) {
        // as a sharded call schedules the row pass: the first seg_top and last seg_bot rows in one
        // segmented launch, then the rows between
        StreamParams q = p;
        if (seg_bot > 0) { q.seg_a = seg_top; q.seg_b = seg_bot; q.seg_b_line0 = d->src_h - seg_bot; }
        else q.n_lines = seg_top;
        if (!emul_dispatch<false>(h.chain, variant, q, warps_h, 0)) return -4;
        const size_t esz = (d->in_type == AVIRB200_U8) ? 1 : (d->in_type == AVIRB200_U16 ? 2 : 4);
        q = p;
        q.n_lines = d->src_h - seg_top - seg_bot;
        q.src = static_cast<const unsigned char*>(src) + (size_t)seg_top * src_pitch * esz;
        q.dst = mid.data() + (size_t)seg_top * d->dst_w * 4;
        if (!emul_dispatch<false>(h.chain, variant, q, warps_h, 0)) return -4;
    } else if (!emul_dispatch<false>(h.chain, variant, p, warps_h, 0)) return -4;

    const int epi = stream_epilogue_code(*d);
    for (int b = 0; b < bands; ++b) {
        stream_fill_params(p, v, *d);
        p.n_lines = d->dst_w;
        p.out0 = (int)((long long)d->dst_h * b / bands);
        p.out1 = (int)((long long)d->dst_h * (b + 1) / bands);
        if (p.out1 <= p.out0) continue;
        p.src = mid.data();
        p.src_pitch = (long long)d->dst_w * 4;
        p.src_row_base = 0;
        g_lo = reinterpret_cast<const unsigned char*>(mid.data());
        g_hi = g_lo + mid.size() * sizeof(float);
        std::vector<float> bandbuf;
        if (need != nullptr) { // the band's own rows only, as on a rank of the sharded schedule
            const int r0 = need[2 * b], nr = need[2 * b + 1];
            const size_t rowf = (size_t)d->dst_w * 4;
            bandbuf.assign(mid.begin() + (size_t)r0 * rowf, mid.begin() + (size_t)(r0 + nr) * rowf);
            p.src = bandbuf.data();
            p.src_row_base = r0;
            p.src_lo = r0;
            p.src_hi = r0 + nr;
            g_lo = reinterpret_cast<const unsigned char*>(bandbuf.data());
            g_hi = g_lo + bandbuf.size() * sizeof(float);
        }
        p.dst = dst;
        p.dst_pitch = (long long)dst_pitch;
        p.dst_type = d->out_type;
        p.dst_row_base = 0;
        if (!emul_dispatch<true>(v.chain, variant, p, warps_v, epi)) return -4;
    }
    g_lo = g_hi = nullptr;
    return g_oob.load() ? -5 : 0;
}

// The sharded schedule with the FUSED halo exchange (AVIRB200_OPT_OVERLAP_HALO = 3): every band's row
// pass writes its own rows into its own buffer and the rows its neighbours need into their mailboxes
// (counting completed boundary rounds, then raising the neighbour's flag); every band's column pass
// reads its own rows from its buffer and the neighbours' rows in place from its mailbox.
// info: avirb200_shard_info of every band (8 ints each).  Every buffer is NaN-poisoned and exactly as
// large as the rows it is meant to hold; reads are bounds-checked (-5), flags and counters are
// checked after the row passes (-6).  Returns 1, nothing computed, where the product would not run the
// fused sender (a band so short that one 16-line strip holds rows of both neighbours).
int stream_emul_resize_fused(const avirb200_plan_desc* d, const void* src, size_t src_pitch, void* dst,
                             size_t dst_pitch, int warps_h, int warps_v, int bands, int variant, const float* lut,
                             int allow, const int* info) {
    StreamAxisPlan h, v;
    if (!stream_row_source_ok(*d) || !stream_plan_axis(d->h, d->sum_mode, d->channels, h, allow) ||
        !stream_plan_axis(d->v, d->sum_mode, d->channels, v, allow))
        return -4;
    struct Band {
        int src_row0, src_rows, dst_row0, dst_rows, need_row0, need_rows, halo_up, halo_down;
        std::vector<float> own, box_up, box_dn; // box_up: rows from the band above, box_dn: from the band below
        unsigned flags[2] = {0, 0};
        unsigned long long count[2] = {0, 0};
    };
    const size_t rowf = (size_t)d->dst_w * 4;
    const float nanv = __builtin_nanf("");
    std::vector<Band> B(bands);
    for (int b = 0; b < bands; ++b) {
        const int* q = info + 8 * b;
        Band& x = B[b];
        x.src_row0 = q[0]; x.src_rows = q[1]; x.dst_row0 = q[2]; x.dst_rows = q[3];
        x.need_row0 = q[4]; x.need_rows = q[5]; x.halo_up = q[6]; x.halo_down = q[7];
        x.own.assign((size_t)x.src_rows * rowf, nanv);
        x.box_up.assign((size_t)x.halo_up * rowf, nanv);
        x.box_dn.assign((size_t)x.halo_down * rowf, nanv);
    }
    const size_t esz = (d->in_type == AVIRB200_U8) ? 1 : (d->in_type == AVIRB200_U16 ? 2 : 4);
    const unsigned seq = 5;
    g_oob = 0;
    g_alt_lo[0] = g_alt_lo[1] = g_alt_hi[0] = g_alt_hi[1] = nullptr;
    StreamParams p;
    for (int b = 0; b < bands; ++b) {
        Band& x = B[b];
        stream_fill_params(p, h, *d);
        p.n_lines = x.src_rows;
        p.out0 = 0;
        p.out1 = d->dst_w;
        p.src = static_cast<const unsigned char*>(src) + (size_t)x.src_row0 * src_pitch * esz;
        p.src_type = stream_row_source_code(*d);
        p.srgb_lut = lut;
        p.src_pitch = (long long)src_pitch;
        p.dst = x.own.data();
        p.dst_pitch = (long long)d->dst_w * 4;
        p.dst_type = AVIRB200_F32;
        g_lo = static_cast<const unsigned char*>(p.src);
        g_hi = g_lo + ((size_t)(x.src_rows - 1) * src_pitch + (size_t)d->src_w * 4) * esz;
        const int top = (b > 0) ? B[b - 1].halo_down : 0, bot = (b + 1 < bands) ? B[b + 1].halo_up : 0;
        if (top > x.src_rows || bot > x.src_rows) return -4;
        if (top > 0 && bot > 0 && (top + kLines - 1) / kLines > (x.src_rows - bot) / kLines) return 1; // a strip with rows of both neighbours: the product pushes
        p.xs_seq = seq;
        p.xs_count = x.count;
        if (top > 0) {
            p.xs_up_dst = B[b - 1].box_dn.data();
            p.xs_up_flag = &B[b - 1].flags[1];
            p.xs_top = top;
            p.xs_units[0] = (unsigned long long)((top + kLines - 1) / kLines);
        }
        if (bot > 0) {
            p.xs_dn_dst = B[b + 1].box_up.data();
            p.xs_dn_flag = &B[b + 1].flags[0];
            p.xs_bot0 = x.src_rows - bot;
            p.xs_bot = bot;
            p.xs_units[1] = (unsigned long long)((x.src_rows + kLines - 1) / kLines - p.xs_bot0 / kLines);
        }
        if (!emul_dispatch<false>(h.chain, variant, p, warps_h, 0)) return -4;
        if (x.count[0] != 0 || x.count[1] != 0) return -6;
        if (top > 0 && B[b - 1].flags[1] != seq) return -6;
        if (bot > 0 && B[b + 1].flags[0] != seq) return -6;
    }
    const int epi = stream_epilogue_code(*d);
    for (int b = 0; b < bands; ++b) {
        Band& x = B[b];
        if (x.dst_rows <= 0) continue;
        stream_fill_params(p, v, *d);
        p.n_lines = d->dst_w;
        p.out0 = x.dst_row0;
        p.out1 = x.dst_row0 + x.dst_rows;
        p.src = x.own.data() - (ptrdiff_t)x.halo_up * (ptrdiff_t)rowf; // (row need_row0 of a buffer that only holds the own rows)
        p.src_pitch = (long long)d->dst_w * 4;
        p.src_row_base = x.need_row0;
        p.src_lo = x.need_row0;
        p.src_hi = x.need_row0 + x.need_rows;
        p.xr_up_src = x.box_up.data();
        p.xr_dn_src = x.box_dn.data();
        p.xr_flags = x.flags;
        p.xr_seq = seq;
        p.xr_own_lo = x.src_row0;
        p.xr_own_hi = x.src_row0 + x.src_rows;
        g_lo = reinterpret_cast<const unsigned char*>(x.own.data());
        g_hi = g_lo + x.own.size() * sizeof(float);
        g_alt_lo[0] = reinterpret_cast<const unsigned char*>(x.box_up.data());
        g_alt_hi[0] = g_alt_lo[0] + x.box_up.size() * sizeof(float);
        g_alt_lo[1] = reinterpret_cast<const unsigned char*>(x.box_dn.data());
        g_alt_hi[1] = g_alt_lo[1] + x.box_dn.size() * sizeof(float);
        if (x.box_up.empty()) g_alt_lo[0] = g_alt_hi[0] = nullptr;
        if (x.box_dn.empty()) g_alt_lo[1] = g_alt_hi[1] = nullptr;
        p.dst = dst;
        p.dst_pitch = (long long)dst_pitch;
        p.dst_type = d->out_type;
        p.dst_row_base = 0;
        if (!emul_dispatch<true>(v.chain, variant, p, warps_v, epi)) return -4;
    }
    g_lo = g_hi = nullptr;
    g_alt_lo[0] = g_alt_lo[1] = g_alt_hi[0] = g_alt_hi[1] = nullptr;
    return g_oob.load() ? -5 : 0;
}

int stream_emul_variants(void) { return kStreamVariants; }

// What the engine would choose for the descriptor: out[0] / out[1] = StreamChainId of the row /
// column pass (0 = the pass runs on the tile or generic kernel), out[2] = row-pass source code,
// out[3] = column-pass epilogue code.  all_chains: as with AVIRB200_STREAM_ALL=1.
void stream_emul_selection(const avirb200_plan_desc* d, int all_chains, int* out) {
    StreamAxisPlan h, v;
    out[0] = (stream_row_source_ok(*d) && stream_plan_axis(d->h, d->sum_mode, d->channels, h, all_chains, false))
                 ? h.chain : 0;
    out[1] = stream_plan_axis(d->v, d->sum_mode, d->channels, v, all_chains, true) ? v.chain : 0;
    out[2] = stream_row_source_code(*d);
    out[3] = stream_epilogue_code(*d);
}

} // extern "C"
