// engine.cu -- libavirb200.so: C ABI (include/avirb200.h) over the sm_100a kernels.
//
// Host responsibilities here are strictly device plumbing: copy the planner's tables into
// one device arena, pick tile sizes that fit shared memory, launch the row pass and the
// column pass, move host images for the convenience entry point, and exchange halo rows
// between row-sharded GPUs.  All arithmetic lives in the kernels.
//
// There is no CPU execution path in this library: without a usable CUDA device every
// entry point fails with AVIRB200_ERR_NO_DEVICE / AVIRB200_ERR_CUDA.

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "avirb200.h"
#include "device_plan.h"
#include "fast_pass.cuh"
#include "generic_pass.cuh"
#include "stream_types.h"
#include "stream_launch.h"

using namespace avb;

namespace {

thread_local std::string g_err;
// debug/test switch: 0 = streaming kernel, else tile kernel, else generic kernel (product
// order); 1 = generic kernel only; 2 = tile kernel, else generic (no streaming kernel)
int g_kernel_mode = 0;
#define g_force_generic (g_kernel_mode == 1)

bool env_stream_enabled() {
    static const bool on = [] {
        const char* e = getenv("AVIRB200_DISABLE_STREAM");
        return !(e && e[0] == '1');
    }();
    return on;
}

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(e_ == cudaErrorMemoryAllocation ? AVIRB200_ERR_ALLOC                 \
                                                        : (e_ == cudaErrorNoDevice ||        \
                                                           e_ == cudaErrorInsufficientDriver \
                                                               ? AVIRB200_ERR_NO_DEVICE      \
                                                               : AVIRB200_ERR_CUDA),         \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

size_t dtype_size(int t) { return t == AVIRB200_U8 ? 1 : (t == AVIRB200_U16 ? 2 : (t == AVIRB200_F64 ? 8 : 4)); }

// The u8 sRGB->linear table: upstream ships 256 float literals (avir.h:234-286) that equal
// the double-precision linearisation formula printed with 7 significant digits.  Regenerated
// here (locale-independent) instead of being copied; tests compare it with the oracle.
void make_srgb_lut(float* lut) {
    for (int i = 0; i < 256; ++i) {
        const double sv = i / 255.0;
        double r;
        if (sv <= 0.04045) {
            r = sv / 12.92;
        } else {
            const double x = (sv + 0.055) / 1.055;
            const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
            r = 0.0985766365536824 + 0.839474952656502 * x2 + 0.363287814061725 * x3 -
                0.0125559718896615 / (0.12758338921578 + 0.290283465468235 * x) -
                0.231757513261358 * x - 0.0395365717969074 * x4;
        }
        char buf[64];
        auto res = std::to_chars(buf, buf + sizeof buf, r, std::chars_format::general, 7);
        float f = 0.0f;
        std::from_chars(buf, res.ptr, f);
        lut[i] = f;
    }
}

struct PassConfig {
    int lines_per_block = 0;
    int tile_out = 0;
    int span = 0;
    int pitch = 0;
    size_t smem = 0;
};

struct HostAxis {
    avirb200_axis_desc desc;              // pointers are HOST copies (below)
    std::vector<std::vector<float> > taps, frac, pdc, sdc;
    std::vector<std::vector<int32_t> > src_pos, phase;
    DevAxis dev;                          // device pointers
    DevAxis hostdev;                      // same geometry, host src_pos pointers (range math)
};

} // namespace

struct avirb200_plan {
    avirb200_plan_desc desc;     // in_type / out_type: what the KERNELS read and write (F64 -> F32)
    int io_in_type = 0, io_out_type = 0; // the caller's element types
    bool errd = false;                   // integer output through the error-diffusion ditherer
    HostAxis h, v;
    void* arena = nullptr;
    float* d_lut = nullptr;
    int device = 0;
    PassConfig cfg_h, cfg_v;
    FastPlan fast;
    avs::StreamAxisPlan stream_h, stream_v; // chain != 0: the pass runs on the streaming kernel
    // resize_host cache
    std::mutex mx;
    void* d_src = nullptr;
    void* d_dst = nullptr;
    void* d_ws = nullptr;
    size_t d_src_bytes = 0, d_dst_bytes = 0, d_ws_bytes = 0;
    cudaStream_t stream = nullptr;
    // pipelined resize_host: copy-in / copy-out streams and per-band events
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_out;
    mutable int last_launches = 0;
};

namespace {

int copy_axis_host(HostAxis& ha, const avirb200_axis_desc& ad) {
    if (ad.nsteps < 1 || ad.nsteps > AVIRB200_MAX_STEPS)
        return fail(AVIRB200_ERR_BAD_ARG, "axis: nsteps out of range");
    ha.desc = ad;
    const int n = ad.nsteps;
    ha.taps.resize(n); ha.frac.resize(n); ha.pdc.resize(n); ha.sdc.resize(n);
    ha.src_pos.resize(n); ha.phase.resize(n);
    int prev_len = ad.src_len;
    int prev_lo = 0, prev_hi = ad.src_len;
    for (int i = 0; i < n; ++i) {
        const avirb200_step_desc& s = ad.steps[i];
        if (s.in_len != prev_len && !(s.kind == AVIRB200_STEP_RESIZE && s.upsampled))
            return fail(AVIRB200_ERR_BAD_ARG, "axis: step in_len does not chain");
        if (s.kind == AVIRB200_STEP_RESIZE && s.upsampled && s.in_len != prev_len)
            return fail(AVIRB200_ERR_BAD_ARG, "axis: upsampled resize in_len does not chain");
        size_t nt = 0;
        if (s.kind == AVIRB200_STEP_RESIZE) {
            if (s.ntaps < 2 || (s.ntaps & 1) || s.nphases < 1 || s.order < 0 || s.order > 1)
                return fail(AVIRB200_ERR_BAD_ARG, "resize step: bad bank geometry");
            nt = (size_t)s.nphases * s.ntaps * (s.order + 1);
            ha.src_pos[i].assign(s.src_pos, s.src_pos + s.out_len);
            ha.phase[i].assign(s.phase, s.phase + s.out_len);
            ha.frac[i].assign(s.frac, s.frac + s.out_len);
            for (int j = 0; j < s.out_len; ++j) {
                if (s.phase[j] < 0 || s.phase[j] >= s.nphases)
                    return fail(AVIRB200_ERR_BAD_ARG, "resize step: phase index out of range");
                if (j > 0 && s.src_pos[j] < s.src_pos[j - 1])
                    return fail(AVIRB200_ERR_BAD_ARG, "resize step: positions not monotonic");
            }
        } else {
            if (s.ntaps < 1) return fail(AVIRB200_ERR_BAD_ARG, "filter step: no taps");
            if (s.kind == AVIRB200_STEP_FIR && s.resample < 1)
                return fail(AVIRB200_ERR_BAD_ARG, "FIR step: resample < 1");
            nt = (size_t)s.ntaps;
            if (s.kind == AVIRB200_STEP_UPSAMPLE) {
                if (s.resample != 2)
                    return fail(AVIRB200_ERR_UNSUPPORTED, "upsample factor other than 2");
                ha.pdc[i].assign(s.prefix_dc, s.prefix_dc + s.n_prefix_dc);
                ha.sdc[i].assign(s.suffix_dc, s.suffix_dc + s.n_suffix_dc);
            }
        }
        ha.taps[i].assign(s.taps, s.taps + nt);
        (void)prev_lo; (void)prev_hi;
        prev_len = s.out_len;
    }
    if (prev_len != ad.dst_len) return fail(AVIRB200_ERR_BAD_ARG, "axis: chain does not end at dst_len");
    return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t axis_arena_bytes(const HostAxis& ha) {
    size_t b = 0;
    for (int i = 0; i < ha.desc.nsteps; ++i) {
        b += align_up(ha.taps[i].size() * 4, 256) + align_up(ha.frac[i].size() * 4, 256) +
             align_up(ha.src_pos[i].size() * 4, 256) + align_up(ha.phase[i].size() * 4, 256) +
             align_up(ha.pdc[i].size() * 4, 256) + align_up(ha.sdc[i].size() * 4, 256);
    }
    return b;
}

template <class T>
const T* stage(std::vector<char>& img, size_t& off, char* dbase, const std::vector<T>& v) {
    if (v.empty()) return nullptr;
    const size_t bytes = v.size() * sizeof(T);
    std::memcpy(img.data() + off, v.data(), bytes);
    const T* d = reinterpret_cast<const T*>(dbase + off);
    off += align_up(bytes, 256);
    return d;
}

void build_dev_axis(HostAxis& ha, std::vector<char>& img, size_t& off, char* dbase) {
    DevAxis& d = ha.dev;
    d.src_len = ha.desc.src_len;
    d.dst_len = ha.desc.dst_len;
    d.nsteps = ha.desc.nsteps;
    int lo = 0, hi = ha.desc.src_len;
    for (int i = 0; i < d.nsteps; ++i) {
        const avirb200_step_desc& s = ha.desc.steps[i];
        DevStep& ds = d.steps[i];
        ds.kind = s.kind; ds.resample = s.resample; ds.latency = s.latency; ds.edge = s.edge;
        ds.in_len = s.in_len; ds.out_len = s.out_len; ds.ntaps = s.ntaps; ds.order = s.order;
        ds.upsampled = s.upsampled; ds.skip_odd = s.skip_odd; ds.zero_start = s.zero_start;
        ds.nphases = s.nphases;
        ds.out_prefix = s.out_prefix; ds.out_suffix = s.out_suffix;
        ds.in_prefix = s.in_prefix; ds.in_suffix = s.in_suffix;
        ds.n_prefix_dc = s.n_prefix_dc; ds.n_suffix_dc = s.n_suffix_dc;
        ds.in_lo = lo; ds.in_hi = hi;
        ds.taps = stage(img, off, dbase, ha.taps[i]);
        ds.src_pos = stage(img, off, dbase, ha.src_pos[i]);
        ds.phase = stage(img, off, dbase, ha.phase[i]);
        ds.frac = stage(img, off, dbase, ha.frac[i]);
        ds.prefix_dc = stage(img, off, dbase, ha.pdc[i]);
        ds.suffix_dc = stage(img, off, dbase, ha.sdc[i]);
        const Range od = step_output_domain(ds);
        lo = od.a;
        hi = od.b + 1;
    }
    ha.hostdev = d;
    for (int i = 0; i < d.nsteps; ++i) {
        ha.hostdev.steps[i].src_pos = ha.src_pos[i].empty() ? nullptr : ha.src_pos[i].data();
        ha.hostdev.steps[i].taps = ha.taps[i].data();
        ha.hostdev.steps[i].phase = ha.phase[i].empty() ? nullptr : ha.phase[i].data();
        ha.hostdev.steps[i].frac = ha.frac[i].empty() ? nullptr : ha.frac[i].data();
    }
}

// Source range a final-output range needs, through the whole chain (host side).
Range chain_source_range(const DevAxis& hd, Range out, int* max_span) {
    Range r = out;
    int span = r.b - r.a + 1;
    for (int i = hd.nsteps - 1; i >= 0; --i) {
        r = step_input_range(hd.steps[i], r, hd.steps[i].src_pos);
        span = imax(span, r.b - r.a + 1);
    }
    if (max_span) *max_span = span;
    return r;
}

const size_t kGenericSmemBudget = 100 * 1024;

PassConfig choose_generic_config(const DevAxis& hd, int channels, int out0, int out1) {
    PassConfig c;
    c.lines_per_block = imax(1, 64 / channels);
    c.pitch = (c.lines_per_block * channels) | 1;
    static const int cand[] = {1024, 768, 512, 384, 256, 192, 128, 96, 64, 48, 32, 24, 16, 12, 8, 4, 2, 1};
    for (int t : cand) {
        int worst = 0;
        for (int j0 = out0; j0 < out1; j0 += t) {
            int sp = 0;
            Range o{j0, imin(j0 + t, out1) - 1};
            chain_source_range(hd, o, &sp);
            worst = imax(worst, sp);
        }
        const size_t smem = 2ull * worst * c.pitch * sizeof(float);
        if (smem <= kGenericSmemBudget || t == 1) {
            c.tile_out = t;
            c.span = worst;
            c.smem = smem;
            break;
        }
    }
    return c;
}

void fill_common(PassParams& p, const avirb200_plan* pl) {
    const avirb200_plan_desc& d = pl->desc;
    p.sum_mode = d.sum_mode;
    p.channels = d.channels;
    p.gamma_in = (d.use_gamma & 1) ? 1 : 0;
    p.gamma_out = (d.use_gamma & 2) ? 1 : 0;
    p.alpha_index = d.alpha_index;
    p.in_gamma_mult = d.in_gamma_mult;
    p.out_gamma_mult = d.out_gamma_mult;
    p.srgb_lut = pl->d_lut;
    p.round_mode = d.round_mode;
    p.tr_mul = d.tr_mul;
    p.tr_mul_inv = d.tr_mul_inv;
    p.pk_out = d.pk_out;
}

int launch_generic(const PassParams& p, const PassConfig& c, cudaStream_t st) {
    dim3 grid((p.out1 - p.out0 + c.tile_out - 1) / c.tile_out,
              (p.n_lines + c.lines_per_block - 1) / c.lines_per_block);
    if (grid.x == 0 || grid.y == 0) return 0;
    if (grid.y > 65535) return fail(AVIRB200_ERR_UNSUPPORTED, "image too large for generic grid");
    if (p.sum_mode == AVIRB200_SUM_DIL8) {
        CUDA_TRY(cudaFuncSetAttribute(generic_pass_kernel<AVIRB200_SUM_DIL8>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        generic_pass_kernel<AVIRB200_SUM_DIL8><<<grid, 256, c.smem, st>>>(p);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(generic_pass_kernel<AVIRB200_SUM_INL>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        generic_pass_kernel<AVIRB200_SUM_INL><<<grid, 256, c.smem, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Row pass over `rows` source rows (band starting at d_src) into the intermediate band
// starting at d_mid; column pass producing dst rows [out0, out1) from an intermediate
// buffer whose row 0 is global row mid_row_base.
int run_row_pass(const avirb200_plan* pl, const void* d_src, size_t src_pitch, float* d_mid,
                 int rows, cudaStream_t st, int* launches) {
    if (rows <= 0) return 0;
    const avirb200_plan_desc& d = pl->desc;
    if (g_kernel_mode == 0 && env_stream_enabled() && pl->stream_h.chain != 0 &&
        ((uintptr_t)d_src % (4 * dtype_size(d.in_type))) == 0 && (src_pitch % 4) == 0 &&
        ((uintptr_t)d_mid % 16) == 0) {
        // (every pixel of the source must be aligned to its own size: the copies move whole pixels)
        avs::StreamParams sp;
        avs::stream_fill_params(sp, pl->stream_h, d);
        sp.src_type = avs::stream_row_source_code(d);
        sp.srgb_lut = pl->d_lut;
        sp.n_lines = rows;
        sp.out0 = 0;
        sp.out1 = d.dst_w;
        sp.src = d_src;
        sp.src_pitch = (long long)src_pitch;
        sp.dst = d_mid;
        sp.dst_pitch = (long long)d.dst_w * 4;
        sp.dst_type = AVIRB200_F32;
        const int r = avs::stream_launch(pl->stream_h.chain, false, 0, sp, st);
        if (r == -1) return fail(AVIRB200_ERR_CUDA, "streaming row pass launch failed");
        if (r == 0) { ++*launches; return 0; }
    }
    if (env_fast_enabled() && !g_force_generic && pl->fast.h_ok) {
        const int r = fast_row_pass(pl->fast, d, d_src, src_pitch, d_mid, rows, pl->d_lut, st);
        if (r == -1) return fail(AVIRB200_ERR_CUDA, "fast row pass launch failed");
        if (r == 0) { ++*launches; return 0; }
    }
    PassParams p;
    std::memset(&p, 0, sizeof p);
    fill_common(p, pl);
    p.ax = pl->h.dev;
    p.is_v = 0;
    p.n_lines = rows;
    p.lines_per_block = pl->cfg_h.lines_per_block;
    p.tile_out = pl->cfg_h.tile_out;
    p.out0 = 0;
    p.out1 = d.dst_w;
    p.span = pl->cfg_h.span;
    p.pitch = pl->cfg_h.pitch;
    p.src = d_src;
    p.src_pitch = (long long)src_pitch;
    p.src_type = d.in_type;
    p.dst = d_mid;
    p.dst_pitch = (long long)d.dst_w * d.channels;
    p.dst_type = AVIRB200_F32;
    ++*launches;
    return launch_generic(p, pl->cfg_h, st);
}

int run_col_pass(const avirb200_plan* pl, const float* d_mid, int mid_row_base, void* d_dst,
                 size_t dst_pitch, int out0, int out1, cudaStream_t st, int* launches) {
    if (out1 <= out0) return 0;
    const avirb200_plan_desc& d = pl->desc;
    {
        const size_t es = fast_elsize(d.out_type);
        if (g_kernel_mode == 0 && env_stream_enabled() && pl->stream_v.chain != 0 &&
            ((uintptr_t)d_dst % (2 * es)) == 0 && (dst_pitch % 2) == 0 && ((uintptr_t)d_mid % 16) == 0) {
            avs::StreamParams sp;
            avs::stream_fill_params(sp, pl->stream_v, d);
            sp.n_lines = d.dst_w;
            sp.out0 = out0;
            sp.out1 = out1;
            sp.src = d_mid;
            sp.src_pitch = (long long)d.dst_w * 4;
            sp.src_row_base = mid_row_base;
            sp.dst = d_dst;
            sp.dst_pitch = (long long)dst_pitch;
            sp.dst_type = d.out_type;
            sp.dst_row_base = out0;
            const int r = avs::stream_launch(pl->stream_v.chain, true, avs::stream_epilogue_code(d), sp, st);
            if (r == -1) return fail(AVIRB200_ERR_CUDA, "streaming column pass launch failed");
            if (r == 0) { ++*launches; return 0; }
        }
    }
    if (env_fast_enabled() && !g_force_generic && pl->fast.v_ok) {
        const int r = fast_col_pass(pl->fast, d, d_mid, mid_row_base, d_dst, dst_pitch, out0, out1,
                                    pl->d_lut, st);
        if (r == -1) return fail(AVIRB200_ERR_CUDA, "fast column pass launch failed");
        if (r == 0) { ++*launches; return 0; }
    }
    PassParams p;
    std::memset(&p, 0, sizeof p);
    fill_common(p, pl);
    p.ax = pl->v.dev;
    p.is_v = 1;
    p.n_lines = d.dst_w;
    PassConfig c = pl->cfg_v;
    if (out0 != 0 || out1 != d.dst_h) c = choose_generic_config(pl->v.hostdev, d.channels, out0, out1);
    p.lines_per_block = c.lines_per_block;
    p.tile_out = c.tile_out;
    p.out0 = out0;
    p.out1 = out1;
    p.span = c.span;
    p.pitch = c.pitch;
    p.src = d_mid;
    p.src_pitch = (long long)d.dst_w * d.channels;
    p.src_type = AVIRB200_F32;
    p.src_row_base = mid_row_base;
    p.dst = d_dst;
    p.dst_pitch = (long long)dst_pitch;
    p.dst_type = d.out_type;
    p.dst_row_base = out0;
    ++*launches;
    return launch_generic(p, c, st);
}

// ---- NCCL through dlopen (no link-time dependency) -------------------------------------------

struct Id128 { char b[128]; }; // ncclUniqueId (passed by value)

struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

Nccl* nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) return;
        n.GetUniqueId = (int (*)(void*))dlsym(n.lib, "ncclGetUniqueId");
        n.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(n.lib, "ncclCommInitRank");
        n.CommDestroy = (int (*)(void*))dlsym(n.lib, "ncclCommDestroy");
        n.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclSend");
        n.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclRecv");
        n.GroupStart = (int (*)())dlsym(n.lib, "ncclGroupStart");
        n.GroupEnd = (int (*)())dlsym(n.lib, "ncclGroupEnd");
        n.GetErrorString = (const char* (*)(int))dlsym(n.lib, "ncclGetErrorString");
    });
    if (!n.lib || !n.GetUniqueId || !n.CommInitRank || !n.Send || !n.Recv || !n.GroupStart ||
        !n.GroupEnd)
        return nullptr;
    return &n;
}

#define NCCL_TRY(expr)                                                                  \
    do {                                                                                \
        int r_ = (expr);                                                                \
        if (r_ != 0)                                                                    \
            return fail(AVIRB200_ERR_NCCL, std::string(#expr) + ": " +                  \
                                               (nc->GetErrorString ? nc->GetErrorString(r_) \
                                                                   : "nccl error"));     \
    } while (0)

int shard_compute_axis(const DevAxis& vaxis, int rank, int nranks, avirb200_shard_info* info) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(AVIRB200_ERR_BAD_ARG, "bad rank");
    const int src_h = vaxis.src_len, dst_h = vaxis.dst_len;
    auto src_split = [&](int r) { return (int)((long long)src_h * r / nranks); };
    auto dst_split = [&](int r) { return (int)((long long)dst_h * r / nranks); };
    info->src_row0 = src_split(rank);
    info->src_rows = src_split(rank + 1) - info->src_row0;
    info->dst_row0 = dst_split(rank);
    info->dst_rows = dst_split(rank + 1) - info->dst_row0;
    if (info->dst_rows <= 0 || info->src_rows <= 0)
        return fail(AVIRB200_ERR_UNSUPPORTED, "image has fewer rows than ranks");
    Range need = chain_source_range(vaxis,
                                    Range{info->dst_row0, info->dst_row0 + info->dst_rows - 1},
                                    nullptr);
    // The band always contains the rank's own rows (they are produced locally anyway).
    need.a = imin(need.a, info->src_row0);
    need.b = imax(need.b, info->src_row0 + info->src_rows - 1);
    info->need_row0 = need.a;
    info->need_rows = need.b - need.a + 1;
    info->halo_up = info->src_row0 - need.a;
    info->halo_down = need.b - (info->src_row0 + info->src_rows - 1);
    if (rank > 0 && info->halo_up > src_split(rank) - src_split(rank - 1))
        return fail(AVIRB200_ERR_UNSUPPORTED, "halo exceeds the neighbouring band (too many ranks)");
    if (rank + 1 < nranks && info->halo_down > src_split(rank + 2 > nranks ? nranks : rank + 2) -
                                                   src_split(rank + 1))
        return fail(AVIRB200_ERR_UNSUPPORTED, "halo exceeds the neighbouring band (too many ranks)");
    return 0;
}

int shard_compute(const avirb200_plan* pl, int rank, int nranks, avirb200_shard_info* info) {
    return shard_compute_axis(pl->v.hostdev, rank, nranks, info);
}

// Geometry-only view of an axis descriptor (host pointers), for range arithmetic.
DevAxis host_axis_view(const avirb200_axis_desc& ad) {
    DevAxis d;
    std::memset(&d, 0, sizeof d);
    d.src_len = ad.src_len; d.dst_len = ad.dst_len; d.nsteps = ad.nsteps;
    int lo = 0, hi = ad.src_len;
    for (int i = 0; i < ad.nsteps && i < AVIRB200_MAX_STEPS; ++i) {
        const avirb200_step_desc& s = ad.steps[i];
        DevStep& ds = d.steps[i];
        ds.kind = s.kind; ds.resample = s.resample; ds.latency = s.latency; ds.edge = s.edge;
        ds.in_len = s.in_len; ds.out_len = s.out_len; ds.ntaps = s.ntaps; ds.order = s.order;
        ds.upsampled = s.upsampled; ds.skip_odd = s.skip_odd; ds.zero_start = s.zero_start;
        ds.nphases = s.nphases;
        ds.out_prefix = s.out_prefix; ds.out_suffix = s.out_suffix;
        ds.in_prefix = s.in_prefix; ds.in_suffix = s.in_suffix;
        ds.n_prefix_dc = s.n_prefix_dc; ds.n_suffix_dc = s.n_suffix_dc;
        ds.in_lo = lo; ds.in_hi = hi;
        ds.taps = s.taps; ds.src_pos = s.src_pos; ds.phase = s.phase; ds.frac = s.frac;
        ds.prefix_dc = s.prefix_dc; ds.suffix_dc = s.suffix_dc;
        const Range od = step_output_domain(ds);
        lo = od.a; hi = od.b + 1;
    }
    return d;
}

// ---- error-diffusion ditherer (upstream CImageResizerDithererErrdINL / ErrdDIL) ---------------
// avir.h:4485-4525, avir_dil.h:927-986, driven row by row from resizeImage (avir.h:5046-5064).
// Per channel, pixel j of row y:   R = (v[j] + D_y[j]) [+ 0.364842 * Noise(j-1)];
//   z = round(R * TrMulI) * TrMul;  Noise = R - z;  out = clamp(z, 0, PkOut);
// and the row below adds D_{y+1}[q] = ((0 + 0.063011*Noise(q-1)) + 0.364842*Noise(q)) + 0.207305*Noise(q+1)
// (the order in which upstream's three "+=" reach the element).  The recursion runs along the
// row AND down the rows, so the parallel form is a wavefront: row y+1 may process pixel q once
// row y has finished pixel q+1.  One warp takes 32 consecutive rows as a systolic array -- lane =
// row, lane l works on pixel t - 2l at step t and hands D_{y+1}[q] to lane l+1 by shuffle, one
// step before it is needed; lane 31 hands its values to lane 0 of the next warp (another block)
// through a row of boundary values in global memory plus a progress counter.  Every block is
// resident at once (one warp each); a block only ever waits for the block before it.
// Quirk kept: the de-interleaved class stores a row as consecutive channel planes and runs them
// one after the other, so the "Dith[j-1] +=" of pixel 0 of plane c+1 lands on the last pixel of
// plane c (avir_dil.h:964, rsdj[-1] with j = 0):  D_{y+1}[c][W-1] gains + 0.207305*Noise_{c+1}(0).
struct ErrdParams {
    const float* src;     // [H][W*C] gamma-corrected floats (the column pass's output)
    void* dst;
    long long dst_pitch;  // elements
    int W, H, dst_type, round_mode;
    int planar;           // de-interleaved class: channel c+1's pixel 0 also feeds channel c's last D (see below)
    float tr_mul, tr_mul_inv, pk_out;
    float* boundary;      // [groups][W*C]
    int* progress;        // [groups]: pixels of the group's last row whose D values are published
};

template <int C>
__global__ void __launch_bounds__(32) errd_kernel(const __grid_constant__ ErrdParams p) {
    const int g = blockIdx.x, lane = threadIdx.x;
    const int W = p.W, y = g * 32 + lane;
    const bool rowok = y < p.H;
    const float* row = p.src + (size_t)(rowok ? y : p.H - 1) * W * C;
    const float* bnd_in = p.boundary + (size_t)(g > 0 ? g - 1 : 0) * W * C;
    float* bnd_out = p.boundary + (size_t)g * W * C;
    volatile int* prog_in = p.progress + (g > 0 ? g - 1 : 0);
    volatile int* prog_out = p.progress + g;
    const bool publish = (lane == 31) && ((g + 1) * 32 < p.H);
    int seen = 0; // lane 0: pixels the group above is known to have published
    float nm1[C], c3p[C], part[C], dn[C], v[C], vn[C], n2first[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { nm1[c] = c3p[c] = part[c] = dn[c] = n2first[c] = 0.0f; v[c] = vn[c] = 0.0f; }
    if (lane == 0 && W > 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) vn[c] = __ldg(row + c);
    }
    const int steps = W + 2 * 31 + 1;
    for (int t = 0; t < steps; ++t) {
        const int pix = t - 2 * lane;
        const bool on = rowok && pix >= 0 && pix < W;
        // D of this pixel: from the lane above (finalised there during the previous step) ...
        float din[C];
#pragma unroll
        for (int c = 0; c < C; ++c) din[c] = __shfl_up_sync(0xffffffffu, dn[c], 1);
        // ... or, for the group's first row, from the group above (row 0 of the image: zero)
        if (lane == 0 && on) {
            if (g == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) din[c] = 0.0f;
            } else {
                while (seen <= pix) seen = *prog_in;
                __threadfence();
#pragma unroll
                for (int c = 0; c < C; ++c) din[c] = __ldcg(bnd_in + (size_t)pix * C + c);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = vn[c];
        // the next pixel's input does not depend on the recursion: fetch it now
        if (rowok && pix + 1 >= 0 && pix + 1 < W) {
#pragma unroll
            for (int c = 0; c < C; ++c) vn[c] = __ldg(row + (size_t)(pix + 1) * C + c);
        }
        if (on) {
            float o[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float R = __fadd_rn(v[c], din[c]);
                if (pix > 0) R = __fadd_rn(R, nm1[c]);
                const float z0 = __fmul_rn(avb::round_out(__fmul_rn(R, p.tr_mul_inv), p.round_mode), p.tr_mul);
                const float noise = __fsub_rn(R, z0);
                o[c] = z0 < 0.0f ? 0.0f : (z0 > p.pk_out ? p.pk_out : z0);
                const float n1 = __fmul_rn(noise, 0.364842f);
                const float n2 = __fmul_rn(noise, 0.207305f);
                const float n3 = __fmul_rn(noise, 0.063011f);
                if (pix == 0) n2first[c] = n2;
                dn[c] = __fadd_rn(part[c], n2); // D_{y+1}[pix-1] is complete (unused for pix == 0)
                part[c] = (pix == 0) ? __fadd_rn(0.0f, n1) : __fadd_rn(__fadd_rn(0.0f, c3p[c]), n1);
                c3p[c] = n3;
                nm1[c] = n1;
            }
            const size_t oi = (size_t)y * (size_t)p.dst_pitch + (size_t)pix * C;
            if (p.dst_type == AVIRB200_U8) {
#pragma unroll
                for (int c = 0; c < C; ++c) static_cast<unsigned char*>(p.dst)[oi + c] = (unsigned char)o[c];
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) static_cast<unsigned short*>(p.dst)[oi + c] = (unsigned short)o[c];
            }
        } else if (rowok && pix == W) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                dn[c] = part[c]; // D_{y+1}[W-1]: no pixel to its right ...
                if (p.planar && c + 1 < C) dn[c] = __fadd_rn(dn[c], n2first[c + 1]); // ... but the next plane's pixel 0
            }
        }
        if (publish) {
            const int q = (on && pix >= 1) ? pix - 1 : ((pix == W) ? W - 1 : -1);
            if (q >= 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) __stcg(bnd_out + (size_t)q * C + c, dn[c]);
                __threadfence();
                *prog_out = q + 1;
            }
        }
    }
}

// ---- double image buffers: the casts upstream's pack / unpack perform, as two small kernels ----

__global__ void __launch_bounds__(256) narrow_f64_kernel(const double* __restrict__ src, long long src_pitch,
                                                        float* __restrict__ dst, int row_elems, int rows) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long n = (long long)row_elems * rows;
    if (i >= n) return;
    const int y = (int)(i / row_elems), x = (int)(i - (long long)y * row_elems);
    dst[i] = __double2float_rn(src[(long long)y * src_pitch + x]); // (fptypeatom) ip[c], avir.h:2803-2806
}

__global__ void __launch_bounds__(256) widen_f32_kernel(const float* __restrict__ src, double* __restrict__ dst,
                                                       long long dst_pitch, int row_elems, int rows) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long n = (long long)row_elems * rows;
    if (i >= n) return;
    const int y = (int)(i / row_elems), x = (int)(i - (long long)y * row_elems);
    dst[(long long)y * dst_pitch + x] = (double)src[i]; // (Tout) v[c], avir.h:3168-3171
}

size_t f64_in_bytes(const avirb200_plan* pl) {
    const avirb200_plan_desc& d = pl->desc;
    return pl->io_in_type == AVIRB200_F64 ? align_up((size_t)d.src_w * d.src_h * d.channels * 4, 256) : 0;
}

size_t f64_out_bytes(const avirb200_plan* pl) { // float copy of the destination (double output, error diffusion)
    const avirb200_plan_desc& d = pl->desc;
    return (pl->io_out_type == AVIRB200_F64 || pl->errd) ? align_up((size_t)d.dst_w * d.dst_h * d.channels * 4, 256) : 0;
}

// error diffusion: per 32-row group one row of boundary values + one progress counter
int errd_groups(const avirb200_plan* pl) { return (pl->desc.dst_h + 31) / 32; }
size_t errd_bytes(const avirb200_plan* pl) {
    if (!pl->errd) return 0;
    const avirb200_plan_desc& d = pl->desc;
    return align_up((size_t)errd_groups(pl) * d.dst_w * d.channels * 4, 256) + align_up((size_t)errd_groups(pl) * 4, 256);
}

bool plan_has_f64(const avirb200_plan* pl) { // plans that only run as a whole image through resize_device / _host
    return pl->io_in_type == AVIRB200_F64 || pl->io_out_type == AVIRB200_F64 || pl->errd;
}

} // namespace

extern "C" {

void avirb200_debug_force_generic(int mode) { g_kernel_mode = (mode == 1 || mode == 2) ? mode : 0; }

int avirb200_plan_kernel_paths(const avirb200_plan* pl) {
    if (pl == nullptr) return 0;
    return (pl->stream_h.chain != 0 ? 1 : 0) | (pl->stream_v.chain != 0 ? 2 : 0) |
           (pl->fast.h_ok ? 4 : 0) | (pl->fast.v_ok ? 8 : 0);
}

int avirb200_shard_query_desc(const avirb200_plan_desc* desc, int rank, int nranks,
                              avirb200_shard_info* info) {
    if (desc == nullptr || info == nullptr) return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    if (desc->v.nsteps < 1 || desc->v.nsteps > AVIRB200_MAX_STEPS)
        return fail(AVIRB200_ERR_BAD_ARG, "axis: nsteps out of range");
    return shard_compute_axis(host_axis_view(desc->v), rank, nranks, info);
}

const char* avirb200_status_string(int s) {
    switch (s) {
    case AVIRB200_OK: return "ok";
    case AVIRB200_ERR_BAD_ARG: return "bad argument";
    case AVIRB200_ERR_CUDA: return "CUDA error";
    case AVIRB200_ERR_NCCL: return "NCCL error";
    case AVIRB200_ERR_UNSUPPORTED: return "unsupported configuration";
    case AVIRB200_ERR_NO_DEVICE: return "no usable CUDA device";
    case AVIRB200_ERR_ALLOC: return "out of memory";
    default: return "unknown status";
    }
}

const char* avirb200_last_error(void) { return g_err.c_str(); }

int avirb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int avirb200_plan_create(const avirb200_plan_desc* desc, avirb200_plan** out) {
    if (desc == nullptr || out == nullptr) return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (desc->channels < 1 || desc->channels > 4 || desc->src_w < 1 || desc->src_h < 1 ||
        desc->dst_w < 1 || desc->dst_h < 1)
        return fail(AVIRB200_ERR_BAD_ARG, "bad image geometry");
    if (desc->in_type < 0 || desc->in_type > 3 || desc->out_type < 0 || desc->out_type > 3)
        return fail(AVIRB200_ERR_BAD_ARG, "bad element type");
    if (desc->h.src_len != desc->src_w || desc->h.dst_len != desc->dst_w ||
        desc->v.src_len != desc->src_h || desc->v.dst_len != desc->dst_h)
        return fail(AVIRB200_ERR_BAD_ARG, "axis lengths do not match the image");
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            return fail(AVIRB200_ERR_NO_DEVICE,
                        std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    std::unique_ptr<avirb200_plan> pl(new (std::nothrow) avirb200_plan());
    if (!pl) return fail(AVIRB200_ERR_ALLOC, "host allocation failed");
    pl->desc = *desc;
    pl->io_in_type = desc->in_type;
    pl->io_out_type = desc->out_type;
    if (desc->in_type == AVIRB200_F64) pl->desc.in_type = AVIRB200_F32;   // cast on the device first
    if (desc->out_type == AVIRB200_F64) pl->desc.out_type = AVIRB200_F32; // widened on the device last
    if (desc->dither == 1 && (desc->out_type == AVIRB200_U8 || desc->out_type == AVIRB200_U16)) {
        // the column pass delivers the gamma-corrected float rows; errd_kernel rounds them in row order
        pl->errd = true;
        pl->desc.out_type = AVIRB200_F32;
        // errd_kernel's blocks (one warp per 32 rows) wait for their predecessor: keep all of them
        // resident at once (148 SMs x 32 blocks) instead of relying on in-order block dispatch
        if ((desc->dst_h + 31) / 32 > 4096)
            return fail(AVIRB200_ERR_UNSUPPORTED, "error diffusion: more than 131072 destination rows");
    }
    int r = copy_axis_host(pl->h, desc->h);
    if (r != 0) return r;
    r = copy_axis_host(pl->v, desc->v);
    if (r != 0) return r;
    CUDA_TRY(cudaGetDevice(&pl->device));

    const size_t bytes = axis_arena_bytes(pl->h) + axis_arena_bytes(pl->v) + 1024 + 256;
    CUDA_TRY(cudaMalloc(&pl->arena, bytes));
    std::vector<char> img(bytes, 0);
    size_t off = 0;
    {
        float lut[256];
        make_srgb_lut(lut);
        std::memcpy(img.data(), lut, sizeof lut);
        pl->d_lut = reinterpret_cast<float*>(pl->arena);
        off = 1024;
    }
    build_dev_axis(pl->h, img, off, static_cast<char*>(pl->arena));
    build_dev_axis(pl->v, img, off, static_cast<char*>(pl->arena));
    CUDA_TRY(cudaMemcpy(pl->arena, img.data(), bytes, cudaMemcpyHostToDevice));

    pl->cfg_h = choose_generic_config(pl->h.hostdev, desc->channels, 0, desc->dst_w);
    pl->cfg_v = choose_generic_config(pl->v.hostdev, desc->channels, 0, desc->dst_h);
    // (pl->desc, not *desc: the kernels' element types, see io_in_type / io_out_type)
    fast_plan_init(pl->fast, pl->h.hostdev, pl->v.hostdev, pl->desc);
    const char* up2e = getenv("AVIRB200_STREAM_ALL"); // tuning switch, see stream_plan_axis()
    const bool up2 = up2e && up2e[0] == '1';
    if (avs::stream_row_source_ok(pl->desc))
        avs::stream_plan_axis(desc->h, desc->sum_mode, desc->channels, pl->stream_h, up2);
    avs::stream_plan_axis(desc->v, desc->sum_mode, desc->channels, pl->stream_v, up2);
    *out = pl.release();
    return 0;
}

void avirb200_plan_destroy(avirb200_plan* pl) {
    if (pl == nullptr) return;
    cudaFree(pl->arena);
    cudaFree(pl->d_src);
    cudaFree(pl->d_dst);
    cudaFree(pl->d_ws);
    fast_plan_free(pl->fast);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    if (pl->stream_in) cudaStreamDestroy(pl->stream_in);
    if (pl->stream_out) cudaStreamDestroy(pl->stream_out);
    for (cudaEvent_t e : pl->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : pl->ev_out) cudaEventDestroy(e);
    delete pl;
}

int avirb200_plan_workspace_bytes(const avirb200_plan* pl, size_t* bytes) {
    if (pl == nullptr || bytes == nullptr) return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    const avirb200_plan_desc& d = pl->desc;
    *bytes = align_up((size_t)d.dst_w * d.src_h * d.channels * sizeof(float), 256) + f64_in_bytes(pl) +
             f64_out_bytes(pl) + errd_bytes(pl);
    return 0;
}

int avirb200_plan_last_launches(const avirb200_plan* pl) { return pl ? pl->last_launches : 0; }

int avirb200_resize_device(const avirb200_plan* pl, const void* d_src, size_t src_pitch, void* d_dst,
                           size_t dst_pitch, void* d_ws, void* stream) {
    if (pl == nullptr || d_src == nullptr || d_dst == nullptr || d_ws == nullptr)
        return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    const avirb200_plan_desc& d = pl->desc;
    if (src_pitch < (size_t)d.src_w * d.channels || dst_pitch < (size_t)d.dst_w * d.channels)
        return fail(AVIRB200_ERR_BAD_ARG, "pitch smaller than a row");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int launches = 0;
    // double buffers: float copies live behind the intermediate in the workspace
    char* wsb = static_cast<char*>(d_ws);
    float* in32 = reinterpret_cast<float*>(wsb + align_up((size_t)d.dst_w * d.src_h * d.channels * 4, 256));
    float* out32 = reinterpret_cast<float*>(reinterpret_cast<char*>(in32) + f64_in_bytes(pl));
    const void* ksrc = d_src;
    size_t ksrc_pitch = src_pitch;
    void* kdst = d_dst;
    size_t kdst_pitch = dst_pitch;
    if (pl->io_in_type == AVIRB200_F64) {
        const int re = d.src_w * d.channels;
        const long long n = (long long)re * d.src_h;
        narrow_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const double*>(d_src),
                                                                       (long long)src_pitch, in32, re, d.src_h);
        ++launches;
        ksrc = in32;
        ksrc_pitch = (size_t)re;
    }
    if (pl->io_out_type == AVIRB200_F64 || pl->errd) {
        kdst = out32;
        kdst_pitch = (size_t)d.dst_w * d.channels;
    }
    int r = run_row_pass(pl, ksrc, ksrc_pitch, static_cast<float*>(d_ws), d.src_h, st, &launches);
    if (r != 0) return r;
    r = run_col_pass(pl, static_cast<const float*>(d_ws), 0, kdst, kdst_pitch, 0, d.dst_h, st,
                     &launches);
    if (r == 0 && pl->io_out_type == AVIRB200_F64) {
        const int re = d.dst_w * d.channels;
        const long long n = (long long)re * d.dst_h;
        widen_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out32, static_cast<double*>(d_dst),
                                                                      (long long)dst_pitch, re, d.dst_h);
        ++launches;
        CUDA_TRY(cudaGetLastError());
    }
    if (r == 0 && pl->errd) {
        ErrdParams ep;
        ep.src = out32;
        ep.dst = d_dst;
        ep.dst_pitch = (long long)dst_pitch;
        ep.W = d.dst_w; ep.H = d.dst_h;
        ep.dst_type = pl->io_out_type;
        ep.round_mode = d.round_mode;
        ep.planar = (d.sum_mode == AVIRB200_SUM_DIL8) ? 1 : 0;
        ep.tr_mul = d.tr_mul; ep.tr_mul_inv = d.tr_mul_inv; ep.pk_out = d.pk_out;
        char* eb = reinterpret_cast<char*>(out32) + f64_out_bytes(pl);
        ep.boundary = reinterpret_cast<float*>(eb);
        ep.progress = reinterpret_cast<int*>(eb + align_up((size_t)errd_groups(pl) * d.dst_w * d.channels * 4, 256));
        CUDA_TRY(cudaMemsetAsync(ep.progress, 0, (size_t)errd_groups(pl) * 4, st));
        switch (d.channels) {
        case 1: errd_kernel<1><<<errd_groups(pl), 32, 0, st>>>(ep); break;
        case 2: errd_kernel<2><<<errd_groups(pl), 32, 0, st>>>(ep); break;
        case 3: errd_kernel<3><<<errd_groups(pl), 32, 0, st>>>(ep); break;
        default: errd_kernel<4><<<errd_groups(pl), 32, 0, st>>>(ep); break;
        }
        ++launches;
        CUDA_TRY(cudaGetLastError());
    }
    pl->last_launches = launches;
    return r;
}

int avirb200_row_pass_device(const avirb200_plan* pl, const void* d_src, size_t src_pitch,
                             void* d_ws, void* stream) {
    if (pl == nullptr || d_src == nullptr || d_ws == nullptr)
        return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    if (plan_has_f64(pl)) return fail(AVIRB200_ERR_UNSUPPORTED, "per-pass entry points: no double buffers, no error diffusion");
    int launches = 0;
    return run_row_pass(pl, d_src, src_pitch, static_cast<float*>(d_ws), pl->desc.src_h,
                        static_cast<cudaStream_t>(stream), &launches);
}

int avirb200_col_pass_device(const avirb200_plan* pl, const void* d_ws, void* d_dst,
                             size_t dst_pitch, void* stream) {
    if (pl == nullptr || d_dst == nullptr || d_ws == nullptr)
        return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    if (plan_has_f64(pl)) return fail(AVIRB200_ERR_UNSUPPORTED, "per-pass entry points: no double buffers, no error diffusion");
    int launches = 0;
    return run_col_pass(pl, static_cast<const float*>(d_ws), 0, d_dst, dst_pitch, 0,
                        pl->desc.dst_h, static_cast<cudaStream_t>(stream), &launches);
}

int avirb200_resize_host(avirb200_plan* pl, const void* h_src, size_t src_pitch, void* h_dst,
                         size_t dst_pitch) {
    if (pl == nullptr || h_src == nullptr || h_dst == nullptr)
        return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    const avirb200_plan_desc& d = pl->desc;
    std::lock_guard<std::mutex> lk(pl->mx);
    CUDA_TRY(cudaSetDevice(pl->device));
    const size_t in_row = (size_t)d.src_w * d.channels * dtype_size(pl->io_in_type);
    const size_t out_row = (size_t)d.dst_w * d.channels * dtype_size(pl->io_out_type);
    const size_t in_bytes = in_row * d.src_h, out_bytes = out_row * d.dst_h;
    size_t ws = 0;
    avirb200_plan_workspace_bytes(pl, &ws);
    if (pl->stream == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    if (pl->d_src_bytes < in_bytes) {
        cudaFree(pl->d_src); pl->d_src = nullptr; pl->d_src_bytes = 0;
        CUDA_TRY(cudaMalloc(&pl->d_src, in_bytes));
        pl->d_src_bytes = in_bytes;
    }
    if (pl->d_dst_bytes < out_bytes) {
        cudaFree(pl->d_dst); pl->d_dst = nullptr; pl->d_dst_bytes = 0;
        CUDA_TRY(cudaMalloc(&pl->d_dst, out_bytes));
        pl->d_dst_bytes = out_bytes;
    }
    if (pl->d_ws_bytes < ws) {
        cudaFree(pl->d_ws); pl->d_ws = nullptr; pl->d_ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&pl->d_ws, ws));
        pl->d_ws_bytes = ws;
    }
    const size_t in_el = dtype_size(pl->io_in_type), out_el = dtype_size(pl->io_out_type);
    // Pipelined form for large images: the image is cut into row bands (the multi-GPU band
    // arithmetic, one shared intermediate buffer instead of a halo exchange).  Band b's rows
    // travel host->device on the copy-in stream while the kernels of band b-1 run on the
    // compute stream and band b-2's destination rows travel back on the copy-out stream: the
    // call takes about as long as the larger of the two PCIe directions instead of their sum.
    // The arithmetic does not depend on the banding (tests: 8-band schedule == unsharded bits).
    int nb = (int)(in_bytes >> 25); // bands of >= 32 MiB of source
    if (nb > 16) nb = 16;
    if (const char* e = getenv("AVIRB200_HOST_BANDS")) nb = atoi(e); // test / tuning switch
    {   // an aliased or overlapping destination (upstream allows NewBuf == SrcBuf) must not be
        // written before the whole source has been read
        const char* s0 = static_cast<const char*>(h_src);
        const char* d0 = static_cast<const char*>(h_dst);
        const char* s1 = s0 + ((size_t)(d.src_h - 1) * src_pitch + (size_t)d.src_w * d.channels) * in_el;
        const char* d1 = d0 + ((size_t)(d.dst_h - 1) * dst_pitch + (size_t)d.dst_w * d.channels) * out_el;
        if (s0 < d1 && d0 < s1) nb = 1;
    }
    if (plan_has_f64(pl)) nb = 1; // the casts / the row-recursive ditherer run over the whole image
    std::vector<avirb200_shard_info> si;
    while (nb >= 2) { // fewer bands until every band's column pass needs only its neighbours' rows
        si.assign(nb, avirb200_shard_info());
        bool ok = true;
        for (int b = 0; b < nb && ok; ++b) ok = (shard_compute(pl, b, nb, &si[b]) == 0);
        if (ok) break;
        nb /= 2;
    }
    if (nb >= 2) {
        if (pl->stream_in == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream_in, cudaStreamNonBlocking));
        if (pl->stream_out == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream_out, cudaStreamNonBlocking));
        while ((int)pl->ev_in.size() < nb) {
            cudaEvent_t e0, e1;
            CUDA_TRY(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
            pl->ev_in.push_back(e0);
            CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            pl->ev_out.push_back(e1);
        }
        const size_t rowf = (size_t)d.dst_w * d.channels;
        const size_t dsrc_pitch = (size_t)d.src_w * d.channels, ddst_pitch = rowf;
        int launches = 0;
        for (int b = 0; b < nb; ++b) {
            CUDA_TRY(cudaMemcpy2DAsync(static_cast<char*>(pl->d_src) + (size_t)si[b].src_row0 * in_row, in_row,
                                       static_cast<const char*>(h_src) + (size_t)si[b].src_row0 * src_pitch * in_el,
                                       src_pitch * in_el, in_row, si[b].src_rows, cudaMemcpyHostToDevice,
                                       pl->stream_in));
            CUDA_TRY(cudaEventRecord(pl->ev_in[b], pl->stream_in));
        }
        auto col_band = [&](int b) -> int {
            char* dd = static_cast<char*>(pl->d_dst) + (size_t)si[b].dst_row0 * out_row;
            int r = run_col_pass(pl, static_cast<const float*>(pl->d_ws), 0, dd, ddst_pitch, si[b].dst_row0,
                                 si[b].dst_row0 + si[b].dst_rows, pl->stream, &launches);
            if (r != 0) return r;
            CUDA_TRY(cudaEventRecord(pl->ev_out[b], pl->stream));
            CUDA_TRY(cudaStreamWaitEvent(pl->stream_out, pl->ev_out[b], 0));
            CUDA_TRY(cudaMemcpy2DAsync(static_cast<char*>(h_dst) + (size_t)si[b].dst_row0 * dst_pitch * out_el,
                                       dst_pitch * out_el, dd, out_row, out_row, si[b].dst_rows,
                                       cudaMemcpyDeviceToHost, pl->stream_out));
            return 0;
        };
        for (int b = 0; b < nb; ++b) {
            CUDA_TRY(cudaStreamWaitEvent(pl->stream, pl->ev_in[b], 0));
            int r = run_row_pass(pl, static_cast<const char*>(pl->d_src) + (size_t)si[b].src_row0 * in_row,
                                 dsrc_pitch, static_cast<float*>(pl->d_ws) + (size_t)si[b].src_row0 * rowf,
                                 si[b].src_rows, pl->stream, &launches);
            if (r != 0) return r;
            if (b > 0 && (r = col_band(b - 1)) != 0) return r; // needs rows of bands b-2 .. b only
        }
        int r = col_band(nb - 1);
        if (r != 0) return r;
        pl->last_launches = launches;
        CUDA_TRY(cudaStreamSynchronize(pl->stream_out));
        CUDA_TRY(cudaStreamSynchronize(pl->stream));
        return 0;
    }
    CUDA_TRY(cudaMemcpy2DAsync(pl->d_src, in_row, h_src, src_pitch * in_el, in_row,
                               d.src_h, cudaMemcpyHostToDevice, pl->stream));
    int r = avirb200_resize_device(pl, pl->d_src, (size_t)d.src_w * d.channels, pl->d_dst,
                                   (size_t)d.dst_w * d.channels, pl->d_ws, pl->stream);
    if (r != 0) return r;
    CUDA_TRY(cudaMemcpy2DAsync(h_dst, dst_pitch * out_el, pl->d_dst, out_row,
                               out_row, d.dst_h, cudaMemcpyDeviceToHost, pl->stream));
    CUDA_TRY(cudaStreamSynchronize(pl->stream));
    return 0;
}

// ---- sharded ---------------------------------------------------------------------------------

int avirb200_shard_query(const avirb200_plan* pl, int rank, int nranks, avirb200_shard_info* info) {
    if (pl == nullptr || info == nullptr) return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    return shard_compute(pl, rank, nranks, info);
}

int avirb200_shard_workspace_bytes(const avirb200_plan* pl, int rank, int nranks, size_t* bytes) {
    if (pl == nullptr || bytes == nullptr) return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    avirb200_shard_info si;
    int r = shard_compute(pl, rank, nranks, &si);
    if (r != 0) return r;
    *bytes = (size_t)si.need_rows * pl->desc.dst_w * pl->desc.channels * sizeof(float);
    return 0;
}

int avirb200_comm_unique_id(void* id128) {
    Nccl* nc = nccl();
    if (!nc) return fail(AVIRB200_ERR_NCCL, "libnccl.so.2 not loadable");
    NCCL_TRY(nc->GetUniqueId(id128));
    return 0;
}

int avirb200_comm_create(const void* id128, int rank, int nranks, void** comm_out) {
    Nccl* nc = nccl();
    if (!nc) return fail(AVIRB200_ERR_NCCL, "libnccl.so.2 not loadable");
    Id128 id;
    std::memcpy(&id, id128, sizeof id);
    NCCL_TRY(nc->CommInitRank(comm_out, nranks, id, rank));
    return 0;
}

void avirb200_comm_destroy(void* comm) {
    Nccl* nc = nccl();
    if (nc && nc->CommDestroy && comm) nc->CommDestroy(comm);
}

int avirb200_resize_sharded(const avirb200_plan* pl, void* comm, int rank, int nranks,
                            const void* d_src, size_t src_pitch, void* d_dst, size_t dst_pitch,
                            void* d_ws, void* stream) {
    if (pl == nullptr || d_src == nullptr || d_dst == nullptr || d_ws == nullptr)
        return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    if (plan_has_f64(pl)) return fail(AVIRB200_ERR_UNSUPPORTED, "sharded calls: no double buffers, no error diffusion");
    avirb200_shard_info si;
    int r = shard_compute(pl, rank, nranks, &si);
    if (r != 0) return r;
    const avirb200_plan_desc& d = pl->desc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t rowf = (size_t)d.dst_w * d.channels; // floats per intermediate row
    float* mid = static_cast<float*>(d_ws);
    float* own = mid + (size_t)si.halo_up * rowf;
    int launches = 0;
    r = run_row_pass(pl, d_src, src_pitch, own, si.src_rows, st, &launches);
    if (r != 0) return r;
    if (nranks > 1) {
        if (comm == nullptr) return fail(AVIRB200_ERR_BAD_ARG, "sharded resize needs a communicator");
        Nccl* nc = nccl();
        if (!nc) return fail(AVIRB200_ERR_NCCL, "libnccl.so.2 not loadable");
        // What the neighbours need from this rank is symmetric information: compute theirs.
        avirb200_shard_info up, down;
        if (rank > 0) { r = shard_compute(pl, rank - 1, nranks, &up); if (r != 0) return r; }
        if (rank + 1 < nranks) { r = shard_compute(pl, rank + 1, nranks, &down); if (r != 0) return r; }
        NCCL_TRY(nc->GroupStart());
        if (rank > 0) {
            if (up.halo_down > 0) // my first rows go up
                NCCL_TRY(nc->Send(own, (size_t)up.halo_down * rowf, 7, rank - 1, comm, st));
            if (si.halo_up > 0)
                NCCL_TRY(nc->Recv(mid, (size_t)si.halo_up * rowf, 7, rank - 1, comm, st));
        }
        if (rank + 1 < nranks) {
            if (down.halo_up > 0) // my last rows go down
                NCCL_TRY(nc->Send(own + (size_t)(si.src_rows - down.halo_up) * rowf,
                                  (size_t)down.halo_up * rowf, 7, rank + 1, comm, st));
            if (si.halo_down > 0)
                NCCL_TRY(nc->Recv(own + (size_t)si.src_rows * rowf, (size_t)si.halo_down * rowf, 7,
                                  rank + 1, comm, st));
        }
        NCCL_TRY(nc->GroupEnd());
    }
    r = run_col_pass(pl, mid, si.need_row0, d_dst, dst_pitch, si.dst_row0,
                     si.dst_row0 + si.dst_rows, st, &launches);
    pl->last_launches = launches;
    return r;
}

int avirb200_resize_sharded_local(const avirb200_plan* pl, int nranks, const void* d_src,
                                  size_t src_pitch, void* d_dst, size_t dst_pitch, void* d_ws,
                                  void* stream) {
    if (pl == nullptr || d_src == nullptr || d_dst == nullptr || d_ws == nullptr)
        return fail(AVIRB200_ERR_BAD_ARG, "null argument");
    if (plan_has_f64(pl)) return fail(AVIRB200_ERR_UNSUPPORTED, "sharded calls: no double buffers, no error diffusion");
    const avirb200_plan_desc& d = pl->desc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t rowf = (size_t)d.dst_w * d.channels;
    std::vector<avirb200_shard_info> si(nranks);
    std::vector<float*> mid(nranks);
    float* base = static_cast<float*>(d_ws);
    for (int r = 0; r < nranks; ++r) {
        int e = shard_compute(pl, r, nranks, &si[r]);
        if (e != 0) return e;
        mid[r] = base;
        base += (size_t)si[r].need_rows * rowf;
    }
    int launches = 0;
    const size_t in_el = dtype_size(d.in_type), out_el = dtype_size(d.out_type);
    for (int r = 0; r < nranks; ++r) { // every band's row pass
        float* own = mid[r] + (size_t)si[r].halo_up * rowf;
        const char* src = static_cast<const char*>(d_src) + (size_t)si[r].src_row0 * src_pitch * in_el;
        int e = run_row_pass(pl, src, src_pitch, own, si[r].src_rows, st, &launches);
        if (e != 0) return e;
    }
    for (int r = 0; r < nranks; ++r) { // the "exchange"
        float* own = mid[r] + (size_t)si[r].halo_up * rowf;
        if (r > 0 && si[r].halo_up > 0) {
            const float* nb = mid[r - 1] + (size_t)(si[r - 1].halo_up + si[r - 1].src_rows - si[r].halo_up) * rowf;
            CUDA_TRY(cudaMemcpyAsync(mid[r], nb, (size_t)si[r].halo_up * rowf * 4,
                                     cudaMemcpyDeviceToDevice, st));
        }
        if (r + 1 < nranks && si[r].halo_down > 0) {
            const float* nb = mid[r + 1] + (size_t)si[r + 1].halo_up * rowf;
            CUDA_TRY(cudaMemcpyAsync(own + (size_t)si[r].src_rows * rowf, nb,
                                     (size_t)si[r].halo_down * rowf * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    for (int r = 0; r < nranks; ++r) {
        char* dst = static_cast<char*>(d_dst) + (size_t)si[r].dst_row0 * dst_pitch * out_el;
        int e = run_col_pass(pl, mid[r], si[r].need_row0, dst, dst_pitch, si[r].dst_row0,
                             si[r].dst_row0 + si[r].dst_rows, st, &launches);
        if (e != 0) return e;
    }
    pl->last_launches = launches;
    return 0;
}

} // extern "C"
