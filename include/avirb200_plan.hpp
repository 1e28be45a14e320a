// avirb200_plan.hpp -- host-side resize planner for the B200 AVIR drop-in (header-only, C++17).
//
// Decides, for one axis of one resizeImage() call, the chain of 1-D filtering steps
// (FIR / filterless 2X upsample / fractional-delay resize), designs their
// coefficients in double precision and produces the per-output resize positions.
// The GPU kernels consume the resulting flat `AxisPlan` through the C ABI in avirb200.h.
//
// The *numerical recipe* (which filters, which lengths, which operation order in the
// double-precision design) is upstream AVIR's and must be reproduced bit-for-bit, since
// the product contract is "same output bits as avir::CImageResizer<>".  Each function
// cites the upstream code whose results it must equal.  The code itself is written from
// scratch around value types and std::vector; nothing is shared with upstream.
//
// IMPORTANT: this header must be compiled without floating-point contraction (FMA fusing
// changes the designed coefficients).  The pragmas below enforce that for GCC/Clang.

#ifndef AVIRB200_PLAN_HPP
#define AVIRB200_PLAN_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#if defined(__clang__)
#pragma clang fp contract(off)
#elif defined(__GNUC__)
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
#endif

namespace avirb200 {
namespace plan {

constexpr double kPi = 3.1415926535897932;    // upstream avir.h:101
constexpr double kPiHalf = 1.5707963267948966; // upstream avir.h:104

// Algorithm parameter set; values of the named presets are upstream's published
// tunables (avir.h:2328-2464), data not code.
struct Params {
    double CorrFltAlpha, CorrFltLen, IntFltAlpha, IntFltCutoff, IntFltLen;
    double LPFltAlpha, LPFltBaseLen, LPFltCutoffMult;
    double HBFltAlpha = 1.94609, HBFltCutoff = 0.46437, HBFltLen = 24; // avir.h:2300-2304
};

inline Params params_preset(int id) {
    Params p{};
    switch (id) {
    default: // Def, avir.h:2332-2339
        p.CorrFltAlpha = 0.97946; p.CorrFltLen = 6.4262; p.IntFltAlpha = 6.41341;
        p.IntFltCutoff = 0.7372; p.IntFltLen = 18; p.LPFltAlpha = 4.76449;
        p.LPFltBaseLen = 7.55999999999998; p.LPFltCutoffMult = 0.79285; break;
    case 1: // ULR, avir.h:2357-2364
        p.CorrFltAlpha = 0.95521; p.CorrFltLen = 5.70774; p.IntFltAlpha = 1.00766;
        p.IntFltCutoff = 0.74202; p.IntFltLen = 18; p.LPFltAlpha = 1.6801;
        p.LPFltBaseLen = 6.62; p.LPFltCutoffMult = 0.67821; break;
    case 2: // LR, avir.h:2381-2388
        p.CorrFltAlpha = 1; p.CorrFltLen = 5.865; p.IntFltAlpha = 1.79529;
        p.IntFltCutoff = 0.74325; p.IntFltLen = 18; p.LPFltAlpha = 1.87597;
        p.LPFltBaseLen = 6.89999999999999; p.LPFltCutoffMult = 0.69326; break;
    case 3: // Low, avir.h:2405-2412
        p.CorrFltAlpha = 0.99739; p.CorrFltLen = 6.20326; p.IntFltAlpha = 4.6836;
        p.IntFltCutoff = 0.73879; p.IntFltLen = 18; p.LPFltAlpha = 7.86565;
        p.LPFltBaseLen = 6.91999999999999; p.LPFltCutoffMult = 0.78379; break;
    case 4: // High, avir.h:2430-2437
        p.CorrFltAlpha = 0.97433; p.CorrFltLen = 6.87893; p.IntFltAlpha = 7.74731;
        p.IntFltCutoff = 0.73844; p.IntFltLen = 18; p.LPFltAlpha = 4.8149;
        p.LPFltBaseLen = 8.07999999999996; p.LPFltCutoffMult = 0.79335; break;
    case 5: // Ultra, avir.h:2455-2462
        p.CorrFltAlpha = 0.99705; p.CorrFltLen = 7.42695; p.IntFltAlpha = 1.71985;
        p.IntFltCutoff = 0.7571; p.IntFltLen = 18; p.LPFltAlpha = 6.71313;
        p.LPFltBaseLen = 8.27999999999996; p.LPFltCutoffMult = 0.78413; break;
    }
    return p;
}

// Which upstream fpclass the call mirrors.  Only three numbers of the fpclass reach the
// planner (avir.h:4576-4587, avir_dil.h:1021-1033): elements per pixel, filter-length
// alignment, and the packing mode used by the complexity model.
struct Mirror {
    int fppack;   // 1 (def, float8_dil) or 4 (float4)
    int elalign;  // 1 (def, float4) or 8 (float8_dil)
    int packmode; // 0 interleaved, 1 de-interleaved
};
constexpr Mirror kMirrorDef{1, 1, 0};
constexpr Mirror kMirrorFloat4{4, 1, 0};
constexpr Mirror kMirrorFloat8Dil{1, 8, 1};

// ---------------------------------------------------------------------------------------
// DSP design primitives

// Two-term sine recurrence; must equal avir.h:1015-1034 (CSineGen).
class SineOsc {
public:
    SineOsc(double step, double phase)
        : cur_(std::sin(phase)), prev_(std::sin(phase - step)), coef_(2.0 * std::cos(step)) {}
    double next() {
        const double r = cur_;
        cur_ = coef_ * r - prev_;
        prev_ = r;
        return r;
    }
private:
    double cur_, prev_, coef_;
};

// Right half of the Peaked Cosine window; must equal avir.h:1065-1084.
class PeakedCosineWin {
public:
    PeakedCosineWin(double alpha, double len2)
        : alpha_(alpha), inv_len2_(1.0 / len2), n_(0.0), osc_(kPiHalf / len2, kPi * 0.5) {}
    double next() {
        const double h = std::pow(n_ * inv_len2_, alpha_);
        n_ += 1.0;
        return osc_.next() * (1.0 - h);
    }
private:
    double alpha_, inv_len2_, n_;
    SineOsc osc_;
};

// Windowed-sinc low-pass; results must equal CDSPPeakedCosineLPF (avir.h:1506-1582).
struct LowPass {
    double len2, freq, alpha;
    int half;  // taps on each side of the centre == latency
    int length;
    LowPass(double len2_, double freq_, double alpha_)
        : len2(len2_), freq(freq_), alpha(alpha_),
          half(static_cast<int>(std::ceil(len2_)) - 1), length(2 * half + 1) {}

    // dc_gain <= 0: no normalisation.
    void design(double* out, double dc_gain) const {
        PeakedCosineWin win(alpha, len2);
        SineOsc osc(freq, 0.0);
        osc.next();
        double* const c = out + half;
        c[0] = freq * win.next();
        double sum = c[0];
        for (int t = 1; t <= half; ++t) {
            const double v = osc.next() * win.next() / t;
            c[t] = v;
            c[-t] = v;
            sum += v + v;
        }
        if (dc_gain > 0.0) {
            const double g = dc_gain / sum;
            for (int i = 0; i < length; ++i) out[i] = out[i] * g;
        }
    }
};

// Response of a float FIR at angular frequency th; must equal avir.h:461-503 with
// fltlat == 0 (the only way the planner calls it).
inline void fir_response(const float* taps, int n, double th, double& re_out, double& im_out) {
    const double coef = 2.0 * std::cos(th);
    double c1 = 1.0, s1 = 0.0;
    double c2 = std::cos(-th), s2 = std::sin(-th);
    double re = 0.0, im = 0.0;
    for (int i = 0; i < n; ++i) {
        const double f = static_cast<double>(taps[i]);
        re += c1 * f;
        im += s1 * f;
        double t = c1; c1 = coef * c1 - c2; c2 = t;
        t = s1; s1 = coef * s1 - s2; s2 = t;
    }
    re_out = re;
    im_out = im;
}

// Scales taps so they sum to dc_gain; must equal avir.h:517-541 for T = double.
inline void normalize_dc(double* p, int n, double dc_gain) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += p[i];
    s = dc_gain / s;
    for (int i = 0; i < n; ++i) p[i] = p[i] * s;
}

// Linear-band FIR equaliser used for the correction filter.  Only the configuration the
// planner uses is implemented: linear band spacing starting at 0 Hz.  Results must equal
// CDSPFIREQ::init + buildFilter (avir.h:1137-1304, 1352-1479) for MinFreq == 0,
// IsLogBands == false.
class BandEq {
public:
    BandEq(double sample_rate, double filter_len, int bands, double max_freq, double win_alpha)
        : bands_(bands) {
        z_ = static_cast<int>(std::ceil(filter_len * 0.5));
        zi_ = z_ + (z_ & 1);
        centers_.assign(bands, 0.0);
        std::vector<double> osc(2 * z_);
        for (int i = 0; i < z_; ++i) { osc[2 * i] = 0.0; osc[2 * i + 1] = 1.0; }
        std::vector<double> win(z_);
        {
            PeakedCosineWin w(win_alpha, filter_len * 0.5);
            for (int i = 1; i <= z_; ++i) win[z_ - i] = w.next();
        }
        lin_.assign(static_cast<size_t>(zi_) * (bands + 1), 0.0);
        ramp_.assign(static_cast<size_t>(zi_) * (bands + 1), 0.0);
        const double step = (max_freq - 0.0) / (bands - 1);
        double f = 0.0;
        f = f * 1.0 + step;
        double x1 = 0.0;
        int kb = 0;
        for (int i = 1; i < bands; ++i) {
            const double x2 = f * 2.0 / sample_rate;
            centers_[i] = x2;
            band_kernel(x1, x2, &lin_[static_cast<size_t>(kb) * zi_],
                        &ramp_[static_cast<size_t>(kb) * zi_], osc, win);
            ++kb;
            x1 = x2;
            f = f * 1.0 + step;
        }
        last_virtual_ = (x1 < 1.0);
        if (last_virtual_)
            band_kernel(x1, 1.0, &lin_[static_cast<size_t>(kb) * zi_],
                        &ramp_[static_cast<size_t>(kb) * zi_], osc, win);
    }
    int length() const { return 2 * z_ - 1; }
    int latency() const { return z_ - 1; }

    void build(const double* gains, double* out) const {
        const double* k1 = lin_.data();
        const double* k2 = ramp_.data();
        double x1 = 0.0, y1 = gains[0];
        double x2 = centers_[1], y2 = gains[1];
        {
            const double c = y1 - y2, d = x1 * y2 - x2 * y1;
            for (int i = 0; i < z_; ++i) out[i] = c * k1[i] + d * k2[i];
        }
        k1 += zi_; k2 += zi_; x1 = x2; y1 = y2;
        for (int b = 2; b < bands_; ++b) {
            x2 = centers_[b];
            y2 = gains[b];
            const double c = y1 - y2, d = x1 * y2 - x2 * y1;
            for (int i = 0; i < z_; ++i) out[i] += c * k1[i] + d * k2[i];
            k1 += zi_; k2 += zi_; x1 = x2; y1 = y2;
        }
        if (last_virtual_) {
            const double c = y1 - y2, d = x1 * y2 - y1;
            for (int i = 0; i < z_; ++i) out[i] += c * k1[i] + d * k2[i];
        }
        for (int i = 0; i < z_ - 1; ++i) out[z_ + i] = out[z_ - 2 - i];
    }

private:
    int bands_, z_, zi_;
    bool last_virtual_;
    std::vector<double> centers_, lin_, ramp_;

    // avir.h:1402-1437
    void band_kernel(double x1, double x2, double* k1, double* k2, std::vector<double>& osc,
                     const std::vector<double>& win) const {
        const double inc = kPi * x2;
        const double coef = 2.0 * std::cos(inc);
        double sv = std::sin(inc * (-z_ + 1));
        double cv = std::sin(inc * (-z_ + 1) + kPi * 0.5);
        osc[0] = std::sin(inc * -z_);
        osc[1] = std::sin(inc * -z_ + kPi * 0.5);
        for (int ks = 1; ks < z_; ++ks) {
            const double ps = osc[2 * ks], pc = osc[2 * ks + 1];
            osc[2 * ks] = sv;
            osc[2 * ks + 1] = cv;
            const double x = kPi * (ks - z_);
            const double v0 = win[ks - 1] / ((x1 - x2) * x);
            k1[ks - 1] = (x2 * sv - x1 * ps + (cv - pc) / x) * v0;
            k2[ks - 1] = (sv - ps) * v0;
            sv = coef * sv - osc[2 * ks - 2];
            cv = coef * cv - osc[2 * ks - 1];
        }
        k1[z_ - 1] = (x2 * x2 - x1 * x1) / (x1 - x2) * 0.5;
        k2[z_ - 1] = -1.0;
    }
};

// A designed low-pass kept in double so it can be folded into the interpolation bank
// ("external filter").  Equality is by design parameters only, as upstream's
// CFltBuffer::operator== (avir.h:1624-1628).
struct ExtFilter {
    std::vector<double> taps;
    double len2 = 0.0, freq = 0.0, alpha = 0.0, dc_gain = 0.0;
    bool same_design(const ExtFilter& o) const {
        return len2 == o.len2 && freq == o.freq && alpha == o.alpha && dc_gain == o.dc_gain;
    }
};

// Bank of fractional-delay filters (order 0, or order 1 = c0 + c1*x), phases created on
// demand.  Observable behaviour -- coefficients, which phases count as "already built",
// the init-cost model -- must equal CDSPFracFilterBankLin<float> (avir.h:1647-2100).
class FracBank {
public:
    FracBank() = default;

    // avir.h:1732-1772
    void configure(int frac_count, int order, double base_len, double cutoff, double alpha,
                   const ExtFilter& ext, int len_align) {
        const double wlen2 = 0.5 * base_len * frac_count;
        const double wfreq = kPi * cutoff / frac_count;
        if (order == order_ && wlen2 == wlen2_ && wfreq == wfreq_ && alpha == walpha_ &&
            frac_count == frac_count_ && ext.same_design(ext_)) {
            needs_init_ = false;
            return;
        }
        wlen2_ = wlen2; wfreq_ = wfreq; walpha_ = alpha;
        frac_count_ = frac_count; order_ = order; ext_ = ext;
        const LowPass lp(wlen2_, wfreq_, walpha_);
        src_len_ = (lp.half / frac_count + 1) * 2;
        len_ = src_len_;
        if (!ext_.taps.empty()) len_ += static_cast<int>(ext_.taps.size()) - 1;
        len_ = (len_ + len_align - 1) & ~(len_align - 1);
        stride_ = len_ * (order + 1);
        src_built_ = false;
        needs_init_ = true;
    }

    // Parameter-only clone used while modelling the V axis (avir.h:1668-1691): flags of the
    // source survive as "non-zero == already built".
    void clone_params(const FracBank& s) {
        wlen2_ = s.wlen2_; wfreq_ = s.wfreq_; walpha_ = s.walpha_;
        frac_count_ = s.frac_count_; order_ = s.order_;
        src_len_ = s.src_len_; len_ = s.len_; stride_ = s.stride_;
        src_built_ = false;
        ext_ = s.ext_;
        flags_.assign(s.flags_.size(), 0);
        for (size_t i = 0; i < s.flags_.size(); ++i)
            flags_[i] = static_cast<uint8_t>(s.flags_[i] << 2);
    }

    bool same_design(const FracBank& o) const { // avir.h:1702-1707
        return order_ == o.order_ && wlen2_ == o.wlen2_ && wfreq_ == o.wfreq_ &&
               walpha_ == o.walpha_ && frac_count_ == o.frac_count_ && ext_.same_design(o.ext_);
    }

    int filter_len() const { return len_; }
    int frac_count() const { return frac_count_; }
    int order() const { return order_; }

    // avir.h:1814-1846.  Returns len*(order+1) floats: c0 then c1.
    const float* phase(int i) {
        if (!src_built_) build_source();
        float* const res = &table_[static_cast<size_t>(i) * stride_];
        if ((flags_[i] & 2) == 0) {
            make_phase(i);
            flags_[i] |= 2;
            if (order_ > 0) {
                make_phase(i + 1);
                const float* const nxt = res + stride_;
                float* const c1 = res + len_;
                for (int j = 0; j < len_; ++j) c1[j] = nxt[j] - res[j];
            }
        }
        return res;
    }
    const float* phase_const(int i) const { return &table_[static_cast<size_t>(i) * stride_]; }
    void build_all() { for (int i = 0; i < frac_count_; ++i) phase(i); }

    // avir.h:1895-1929
    int init_cost(const std::vector<char>& used) const {
        const int use_cost = len_ * order_ + src_len_ * static_cast<int>(ext_.taps.size());
        int ic;
        if (needs_init_) {
            ic = frac_count_ * src_len_ * 65;
            for (int i = 0; i < frac_count_; ++i) if (used[i]) ic += use_cost;
        } else {
            ic = 0;
            for (int i = 0; i < frac_count_; ++i)
                if (used[i] != 0 && flags_[i] == 0) ic += use_cost;
        }
        return ic;
    }

private:
    double wlen2_ = 0.0, wfreq_ = 0.0, walpha_ = 0.0;
    int frac_count_ = 0, order_ = -1;
    int src_len_ = 0, len_ = 0, stride_ = 0;
    bool needs_init_ = false, src_built_ = false;
    ExtFilter ext_;
    std::vector<float> table_;
    std::vector<uint8_t> flags_;
    std::vector<double> src_;

    // avir.h:1970-2009: one long low-pass, polyphase-split into frac_count+1 filters,
    // each normalised to unity DC.
    void build_source() {
        src_built_ = true;
        needs_init_ = false;
        const LowPass lp(wlen2_, wfreq_, walpha_);
        const int buf_len = src_len_ * frac_count_ + 1;
        const int centre = src_len_ * frac_count_ / 2;
        std::vector<double> buf(buf_len, 0.0);
        lp.design(&buf[centre - lp.half], 0.0);
        src_.assign(static_cast<size_t>(frac_count_ + 1) * src_len_, 0.0);
        flags_.assign(frac_count_ + 1, 0);
        double* op = src_.data();
        for (int i = frac_count_; i >= 0; --i) {
            const double* ip = buf.data() + i;
            for (int j = 0; j < src_len_; ++j) { op[j] = *ip; ip += frac_count_; }
            normalize_dc(op, src_len_, 1.0);
            op += src_len_;
        }
        table_.assign(static_cast<size_t>(frac_count_ + 1) * stride_, 0.0f);
    }

    // avir.h:2021-2099.  NOTE the source table is stored in reverse phase order: slot n of
    // the table is built from source row n (row 0 holds polyphase branch frac_count).
    void make_phase(int n) {
        if (flags_[n] != 0) return;
        flags_[n] |= 1;
        const int ext_n = static_cast<int>(ext_.taps.size());
        const int res_lat = ext_n / 2 + src_len_ / 2;
        int res_len = src_len_;
        if (ext_n > 0) res_len += ext_n - 1;
        const int offs = len_ / 2 - res_lat;
        float* op = &table_[static_cast<size_t>(n) * stride_];
        for (int i = 0; i < offs; ++i) op[i] = 0.0f;
        for (int i = offs + res_len; i < len_; ++i) op[i] = 0.0f;
        op += offs;
        const double* const sf = &src_[static_cast<size_t>(n) * src_len_];
        if (ext_n == 0) {
            for (int i = 0; i < res_len; ++i) op[i] = static_cast<float>(sf[i]);
            return;
        }
        const double* const ef = ext_.taps.data();
        for (int j = 0; j < res_len; ++j) {
            int k = 0;
            int l = j - ext_n + 1;
            int r = l + ext_n;
            if (l < 0) { k -= l; l = 0; }
            if (r > src_len_) r = src_len_;
            double s = 0.0;
            const int cnt = r - l;
            for (int i = 0; i < cnt; ++i) s += ef[k + i] * sf[l + i];
            op[j] = static_cast<float>(s);
        }
    }
};

// ---------------------------------------------------------------------------------------
// Filtering steps

enum StepKind : int { kStepFir = 0, kStepUpsample = 1, kStepResize = 2 };

struct ResizePos {
    int src_pos; // integer source position (coordinates of this step's input line)
    int fti;     // phase index into the bank
    float x;     // order-1 interpolation fraction
};

// Mutable planning record for one step (fields as upstream's CImageResizerFilterStep,
// avir.h:2568-2728, minus everything only the CPU scanline code needs).
struct Step {
    bool is_upsample = false;
    int resample = 0; // 0 = resize step
    std::vector<float> flt;
    int flt_cap = 0;  // filter capacity (== flt.size() once built; set alone when modelling)
    ExtFilter orig;   // double-precision original, kept when it is folded into the bank
    bool orig_present = false;
    int prefix_dc_cap = 0, suffix_dc_cap = 0;
    std::vector<float> prefix_dc, suffix_dc; // filtered upsample edge tails
    double dc_gain = 1.0;
    int latency = 0;
    int in_len = 0, in_prefix = 0, in_suffix = 0;
    int out_len = 0, out_prefix = 0, out_suffix = 0;
    int edge = 0;
    // resize step
    FracBank* bank = nullptr; // the bank in use (fixed or dynamic)
    bool bank_dynamic = false;
    double pos_k = 0.0, pos_o = 0.0;
    std::vector<ResizePos> pos;
};

struct AxisVars {
    double k = 1.0, o = 0.0;
    int resize_step = 0;
    bool is_resize2 = false;
};

// Immutable per-resizer state: parameters, bit depths, the fixed (k-independent) bank.
// Mirrors what avir::CImageResizer<> holds (avir.h:4630-4639, 5101-5106).
class Designer {
public:
    Designer(int res_bits, int src_bits, const Params& p, const Mirror& m)
        : prm_(p), mir_(m), res_bits_(res_bits), src_bits_(src_bits == 0 ? res_bits : src_bits) {
        init_bank(fixed_, 1.0, false, ExtFilter());
        fixed_.build_all();
    }
    const Mirror& mirror() const { return mir_; }
    const Params& params() const { return prm_; }
    int res_bits() const { return res_bits_; }
    int fixed_order() const { return fixed_.order(); }

    // avir.h:5128-5164
    void init_bank(FracBank& b, double cutoff_mult, bool force_hi, const ExtFilter& ext) const {
        const int bits = res_bits_ > src_bits_ ? res_bits_ : src_bits_;
        const double snr = -6.02 * (bits + 3);
        int order, fc;
        if (force_hi || bits > 8) {
            order = 1;
            fc = static_cast<int>(std::ceil(0.23134052 * std::exp(-0.058062929 * snr)));
        } else {
            order = 0;
            fc = static_cast<int>(std::ceil(0.33287686 * std::exp(-0.11334583 * snr)));
        }
        if (fc < 2) fc = 2;
        b.configure(fc, order, prm_.IntFltLen / cutoff_mult, prm_.IntFltCutoff * cutoff_mult,
                    prm_.IntFltAlpha, ext, mir_.elalign);
    }

    // Rounds a filter capacity up to the mirror's element alignment (avir.h:5181-5207).
    int aligned(int n) const { return (n + mir_.elalign - 1) & ~(mir_.elalign - 1); }

    // avir.h:5231-5360
    void assign_filter(Step& fs, bool is_up, int factor, double cutoff, double dc_gain,
                       bool keep_orig, bool model) const {
        double alpha, len2, freq;
        if (cutoff == 0.0) {
            const double m = 2.0 / factor;
            alpha = prm_.HBFltAlpha;
            len2 = 0.5 * prm_.HBFltLen / m;
            freq = kPi * prm_.HBFltCutoff * m;
        } else {
            alpha = prm_.LPFltAlpha;
            len2 = 0.25 * prm_.LPFltBaseLen / cutoff;
            freq = kPi * prm_.LPFltCutoffMult * cutoff;
        }
        if (is_up) {
            len2 *= factor;
            freq /= factor;
            fs.dc_gain = dc_gain * factor;
        } else {
            fs.dc_gain = dc_gain;
        }
        fs.orig.len2 = len2; fs.orig.freq = freq; fs.orig.alpha = alpha;
        fs.orig.dc_gain = fs.dc_gain;
        const LowPass lp(len2, freq, alpha);
        fs.is_upsample = is_up;
        fs.resample = factor;
        fs.latency = lp.half;
        fs.flt_cap = aligned(lp.length);
        const int ext = fs.flt_cap - lp.length;
        if (model) {
            fs.flt.clear();
            if (keep_orig) { fs.orig.taps.assign(lp.length, 0.0); fs.orig_present = true; }
        } else {
            fs.orig.taps.assign(lp.length, 0.0);
            lp.design(fs.orig.taps.data(), fs.dc_gain);
            fs.flt.assign(fs.flt_cap, 0.0f);
            for (int i = 0; i < lp.length; ++i) fs.flt[i] = static_cast<float>(fs.orig.taps[i]);
            fs.orig_present = keep_orig;
            if (!keep_orig) fs.orig.taps.clear();
        }
        if (is_up) {
            int l = fs.flt_cap - fs.latency - factor - ext;
            fs.prefix_dc_cap = aligned(l);
            fs.suffix_dc_cap = aligned(fs.latency);
            if (!model) {
                // Edge "DC tails" (avir.h:5320-5353): float sums of the taps that fall
                // beyond the explicitly filtered edge replicas.
                fs.prefix_dc.assign(fs.prefix_dc_cap, 0.0f);
                fs.suffix_dc.assign(fs.suffix_dc_cap, 0.0f);
                const float* ip = &fs.flt[fs.latency + factor];
                for (int i = 0; i < l; ++i) fs.prefix_dc[i] = ip[i];
                while (true) {
                    ip += factor;
                    l -= factor;
                    if (l <= 0) break;
                    for (int i = 0; i < l; ++i) fs.prefix_dc[i] += ip[i];
                }
                l = fs.latency;
                float* op = fs.suffix_dc.data();
                for (int i = 0; i < l; ++i) op[i] = fs.flt[i];
                while (true) {
                    op += factor;
                    l -= factor;
                    if (l <= 0) break;
                    for (int i = 0; i < l; ++i) op[i] += fs.flt[i];
                }
            }
        } else if (!keep_orig) {
            fs.edge = 3; // EdgePixelCountDef, avir.h:2629
        }
    }

    // avir.h:5384-5487.  `steps` holds every step built so far; when `pre` the correction
    // filter occupies steps[0] and corrects for steps[1..], otherwise it is appended and
    // corrects for everything before it.
    void add_correction(std::vector<std::unique_ptr<Step>>& steps, double bw, bool pre,
                        bool model) const {
        if (!pre) steps.emplace_back(new Step());
        Step& nfs = pre ? *steps.front() : *steps.back();
        nfs.is_upsample = false;
        nfs.resample = 1;
        nfs.dc_gain = 1.0;
        nfs.edge = pre ? 3 : 0;
        if (model) {
            const int l = static_cast<int>(std::ceil(prm_.CorrFltLen * 0.5));
            nfs.latency = l - 1;
            nfs.flt_cap = aligned(l * 2 - 1);
            nfs.flt.clear();
            return;
        }
        constexpr int kBins = 65;
        double curbw = 1.0;
        double bins[kBins];
        for (double& b : bins) b = 1.0;
        const int si = pre ? 1 : 0;
        const int n = static_cast<int>(steps.size());
        for (int i = si; i < n - (si ^ 1); ++i) {
            Step& fs = *steps[i];
            if (fs.is_upsample) {
                curbw *= fs.resample;
                if (fs.orig_present) continue;
            }
            const float* taps;
            int ntaps;
            if (fs.resample == 0) {
                taps = fs.bank_dynamic ? fs.bank->phase(0) : fs.bank->phase_const(0);
                ntaps = fs.bank->filter_len();
            } else {
                taps = fs.flt.data();
                ntaps = fs.flt_cap;
            }
            const double thm = kPi * bw / (curbw * (kBins - 1));
            for (int j = 0; j < kBins; ++j) {
                double re, im;
                fir_response(taps, ntaps, j * thm, re, im);
                bins[j] *= fs.dc_gain / std::sqrt(re * re + im * im);
            }
            if (!fs.is_upsample && fs.resample > 1) curbw /= fs.resample;
        }
        const BandEq eq(bw * 2.0, prm_.CorrFltLen, kBins, bw, prm_.CorrFltAlpha);
        nfs.latency = eq.latency();
        std::vector<double> f(eq.length());
        eq.build(bins, f.data());
        normalize_dc(f.data(), eq.length(), 1.0);
        nfs.flt_cap = aligned(eq.length());
        nfs.flt.assign(nfs.flt_cap, 0.0f);
        for (int i = 0; i < eq.length(); ++i) nfs.flt[i] = static_cast<float>(f[i]);
    }

    // avir.h:5616-5739
    void build_steps(std::vector<std::unique_ptr<Step>>& steps, AxisVars& v, FracBank& dyn,
                     double dc_gain, int mode, bool model) {
        steps.clear();
        const bool combo = (mode & 1) != 0;
        const bool force_hi = (mode & 2) != 0;
        const bool halfband = (mode & 4) != 0;
        const double bw = 1.0 / v.k;
        const int up = static_cast<int>(std::floor(v.k)) < 2 ? 2 : 1;
        double int_mult, cutoff, corrbw;
        bool pre;
        Step* reuse = nullptr;
        Step* ext_step = nullptr;
        if (v.k <= 1.0) {
            pre = true; cutoff = 1.0; corrbw = 1.0;
            steps.emplace_back(new Step());
        } else {
            pre = false; cutoff = bw; corrbw = bw;
        }
        if (up > 1) {
            steps.emplace_back(new Step());
            Step& fs = *steps.back();
            assign_filter(fs, true, up, cutoff, dc_gain, combo, model);
            int_mult = cutoff * 2.0 / up;
            ext_step = combo ? &fs : nullptr;
        } else {
            int down;
            while (true) {
                down = static_cast<int>(std::floor(0.5 / cutoff));
                if (halfband && down > 1) {
                    steps.emplace_back(new Step());
                    assign_filter(*steps.back(), false, down, 0.0, 1.0, false, model);
                    cutoff *= down;
                } else {
                    if (down < 1) down = 1;
                    break;
                }
            }
            steps.emplace_back(new Step());
            Step& fs = *steps.back();
            assign_filter(fs, false, down, cutoff, dc_gain, combo, model);
            int_mult = cutoff / 0.5;
            if (combo) { reuse = &fs; ext_step = &fs; }
            else int_mult *= down;
        }
        if (reuse == nullptr) steps.emplace_back(new Step());
        Step& rs = reuse ? *reuse : *steps.back();
        v.resize_step = static_cast<int>(steps.size()) - 1;
        rs.is_upsample = false;
        rs.resample = 0;
        rs.dc_gain = ext_step ? ext_step->dc_gain : 1.0;
        {
            static const ExtFilter none;
            const ExtFilter& e = ext_step ? ext_step->orig : rs.orig;
            // A step that does not keep its original exposes an empty external filter
            // (upstream frees FltOrig, avir.h:5301-5304); parameters still compare.
            ExtFilter use = e;
            if (!(ext_step ? ext_step->orig_present : rs.orig_present)) use.taps.clear();
            (void)none;
            init_bank(dyn, int_mult, force_hi, use);
        }
        if (dyn.same_design(fixed_)) { rs.bank = &fixed_; rs.bank_dynamic = false; }
        else { rs.bank = &dyn; rs.bank_dynamic = true; }
        add_correction(steps, corrbw, pre, model);
    }

    // avir.h:5827-5937 (+ fillRPosBuf 5782-5808, extendUpsample 5753-5766).
    void size_steps(std::vector<std::unique_ptr<Step>>& steps, AxisVars& v, int src_len,
                    int new_len) const {
        int upstep = -1;
        const int n = static_cast<int>(steps.size());
        for (int i = 0; i < n; ++i) {
            Step& fs = *steps[i];
            fs.in_len = src_len;
            if (fs.is_upsample) {
                upstep = i;
                v.k *= fs.resample;
                v.o *= fs.resample;
                fs.in_prefix = 0;
                fs.in_suffix = 0;
                fs.out_len = fs.in_len * fs.resample;
                fs.out_prefix = fs.latency;
                fs.out_suffix = fs.flt_cap - fs.latency - fs.resample;
                int l0 = fs.out_prefix + fs.out_len + fs.out_suffix;
                const int l = fs.in_len * fs.resample + fs.suffix_dc_cap;
                if (l > l0) fs.out_suffix += l - l0;
                l0 = fs.out_len + fs.out_suffix;
                if (fs.prefix_dc_cap > l0) fs.out_suffix += fs.prefix_dc_cap - l0;
            } else if (fs.resample == 0) {
                const int half = fs.bank->filter_len() / 2;
                const int lpix = static_cast<int>(std::floor(v.o)) - (half - 1);
                fs.in_prefix = lpix < 0 ? -lpix : 0;
                const int rpix =
                    static_cast<int>(std::floor(v.o + (new_len - 1) * v.k)) + half + 1;
                fs.in_suffix = rpix > fs.in_len ? rpix - fs.in_len : 0;
                fs.out_len = new_len;
                fs.pos_k = v.k;
                fs.pos_o = v.o;
                const int fc = fs.bank->frac_count();
                fs.pos.resize(new_len);
                for (int j = 0; j < new_len; ++j) {
                    const double sp = v.o + v.k * j;
                    const int spi = static_cast<int>(std::floor(sp));
                    const double x = (sp - spi) * fc;
                    const int fti = static_cast<int>(x);
                    fs.pos[j].x = static_cast<float>(x - fti);
                    fs.pos[j].fti = fti;
                    fs.pos[j].src_pos = spi;
                }
            } else {
                v.k /= fs.resample;
                v.o /= fs.resample;
                v.o += fs.edge;
                fs.in_prefix = fs.latency;
                fs.in_suffix = fs.flt_cap - fs.latency - 1;
                fs.out_len = (fs.in_len + fs.resample - 1) / fs.resample + fs.edge;
                fs.in_suffix += (fs.out_len - 1) * fs.resample + 1 - fs.in_len;
                fs.in_prefix += fs.edge * fs.resample;
                fs.out_len += fs.edge;
            }
            src_len = fs.out_len;
        }
        v.is_resize2 = false;
        if (upstep != -1) {
            Step& us = *steps[upstep];
            Step& nx = *steps[upstep + 1];
            us.in_prefix = (nx.in_prefix + us.resample - 1) / us.resample;
            us.out_prefix += us.in_prefix * us.resample;
            nx.in_prefix = 0;
            us.in_suffix = (nx.in_suffix + us.resample - 1) / us.resample;
            us.out_suffix += us.in_suffix * us.resample;
            nx.in_suffix = 0;
            if (us.resample == 2 && v.resize_step == upstep + 1 && mir_.packmode == 0 &&
                us.orig_present)
                v.is_resize2 = true;
        }
    }

    // Makes sure every phase the positions refer to exists (upstream does this as a side
    // effect of updateBufLenAndRPosPtrs, avir.h:6063-6126); the created-phase flags feed
    // the V-axis complexity model.
    static void touch_phases(Step& rs) {
        if (!rs.bank_dynamic) return;
        for (const ResizePos& p : rs.pos) rs.bank->phase(p.fti);
    }

    // avir.h:6206-6270 (+ fillUsedFracMap 6167-6183)
    int complexity(const std::vector<std::unique_ptr<Step>>& steps, const AxisVars& v,
                   int el_count, int scanlines) const {
        const int fcnum = mir_.packmode != 0 ? 1 : 3;
        const int fcden = mir_.packmode != 0 ? 1 : 4;
        int s = 0, s2 = 0;
        const int n = static_cast<int>(steps.size());
        for (int i = 0; i < n; ++i) {
            const Step& fs = *steps[i];
            s2 += 65 * fs.flt_cap;
            if (fs.is_upsample) {
                if (fs.orig_present) continue;
                s += (fs.flt_cap * (fs.in_prefix + fs.in_len + fs.in_suffix) + fs.suffix_dc_cap +
                      fs.prefix_dc_cap) * el_count;
            } else if (fs.resample == 0) {
                s += fs.bank->filter_len() * (fs.bank->order() + el_count) * fs.out_len;
                if (i == v.resize_step && v.is_resize2) s >>= 1;
                std::vector<char> used(fs.bank->frac_count(), 0);
                for (const ResizePos& p : fs.pos) used[p.fti] |= 1;
                s2 += fs.bank->init_cost(used);
            } else {
                s += fs.flt_cap * el_count * fs.out_len * fcnum / fcden;
            }
        }
        return s + s2 / scanlines;
    }

    // avir.h:6137-6157
    static void scale_correction(std::vector<std::unique_ptr<Step>>& steps, double m) {
        Step& last = *steps.back();
        Step& tgt = (!last.is_upsample && last.resample == 1) ? last : *steps.front();
        for (float& f : tgt.flt) f = static_cast<float>(static_cast<double>(f) * m);
    }

    FracBank& fixed_bank() { return fixed_; }

private:
    Params prm_;
    Mirror mir_;
    int res_bits_, src_bits_;
    FracBank fixed_;
};

// ---------------------------------------------------------------------------------------
// Flat, executable per-axis plan (what the kernels need, nothing else)

struct ExecStep {
    int kind = 0;       // StepKind; filterless upsample is folded into the following resize
    int resample = 1;   // FIR: decimation factor R
    int latency = 0;    // FIR: L
    int edge = 0;       // FIR: extra outputs per side
    int in_len = 0;     // length of this step's input line (clamp domain)
    int out_len = 0;
    int ntaps = 0;      // FIR: stored taps (incl. alignment zeros); resize: bank filter length
    int order = 0;      // resize: 0 or 1
    int upsampled = 0;  // resize: input is the virtual 2X zero-stuffed line of `in_len` samples
    int skip_odd = 0;   // resize: upstream's doResize2 (only taps landing on real samples)
    int zero_start = 0; // accumulators start at +0 instead of at the first product
    int out_prefix = 0, out_suffix = 0, in_prefix = 0, in_suffix = 0; // filtered upsample
    std::vector<float> prefix_dc, suffix_dc;                          // filtered upsample
    std::vector<float> taps;      // FIR taps, or bank phases [nphases][ntaps*(order+1)]
    int nphases = 0;
    std::vector<int32_t> src_pos; // resize, per output
    std::vector<int32_t> phase;   // resize, per output: index into `taps` phases
    std::vector<float> frac;      // resize, per output
};

struct AxisPlan {
    int mode = 0;      // build mode chosen
    int src_len = 0, dst_len = 0;
    std::vector<ExecStep> steps;
    bool unsupported = false; // chain contains a filtered upsample (not on the GPU path yet)
};

inline AxisPlan flatten(std::vector<std::unique_ptr<Step>>& steps, const AxisVars& v, int mode,
                        int src_len, int dst_len) {
    AxisPlan ap;
    ap.mode = mode;
    ap.src_len = src_len;
    ap.dst_len = dst_len;
    bool pending_up = false;
    int up_in_len = 0;
    for (size_t i = 0; i < steps.size(); ++i) {
        Step& fs = *steps[i];
        if (fs.is_upsample) {
            if (fs.resample != 2) { ap.unsupported = true; return ap; }
            if (fs.orig_present) { // filterless: folded into the resize step that follows
                pending_up = true;
                up_in_len = fs.in_len;
                continue;
            }
            ExecStep us;
            us.kind = kStepUpsample;
            us.resample = fs.resample;
            us.latency = fs.latency;
            us.in_len = fs.in_len;
            us.out_len = fs.out_len;
            us.ntaps = fs.flt_cap;
            us.taps = fs.flt;
            us.out_prefix = fs.out_prefix; us.out_suffix = fs.out_suffix;
            us.in_prefix = fs.in_prefix; us.in_suffix = fs.in_suffix;
            us.prefix_dc = fs.prefix_dc; us.suffix_dc = fs.suffix_dc;
            ap.steps.push_back(std::move(us));
            continue;
        }
        ExecStep es;
        if (fs.resample == 0) {
            es.kind = kStepResize;
            es.ntaps = fs.bank->filter_len();
            es.order = fs.bank->order();
            es.out_len = fs.out_len;
            if (pending_up) {
                es.upsampled = 1;
                es.skip_odd = v.is_resize2 ? 1 : 0;
                es.in_len = up_in_len;
                pending_up = false;
            } else {
                es.in_len = fs.in_len;
            }
            const int stride = es.ntaps * (es.order + 1);
            std::vector<int> slot(fs.bank->frac_count() + 1, -1);
            es.src_pos.resize(fs.out_len);
            es.phase.resize(fs.out_len);
            es.frac.resize(fs.out_len);
            for (int j = 0; j < fs.out_len; ++j) {
                const ResizePos& p = fs.pos[j];
                if (slot[p.fti] < 0) {
                    slot[p.fti] = es.nphases++;
                    const float* t = fs.bank_dynamic ? fs.bank->phase(p.fti)
                                                     : fs.bank->phase_const(p.fti);
                    es.taps.insert(es.taps.end(), t, t + stride);
                }
                es.src_pos[j] = p.src_pos;
                es.phase[j] = slot[p.fti];
                es.frac[j] = p.x;
            }
        } else {
            if (pending_up) { ap.unsupported = true; return ap; }
            es.kind = kStepFir;
            es.resample = fs.resample;
            es.latency = fs.latency;
            es.edge = fs.edge;
            es.in_len = fs.in_len;
            es.out_len = fs.out_len;
            es.ntaps = fs.flt_cap;
            es.taps = fs.flt;
        }
        ap.steps.push_back(std::move(es));
    }
    return ap;
}

// Everything resizeImage() decides on the host for one call (both axes).
struct ImagePlan {
    AxisPlan h, v;
    double kx = 1, ky = 1, ox = 0, oy = 0;
    double out_mul = 1.0;
    double in_gamma_mult = 0.0, out_gamma_mult = 0.0;
    int el_count = 0;
};

struct CallDesc {
    int src_w, src_h, new_w, new_h, channels;
    double k;          // as passed to resizeImage
    double ox, oy;     // CImageResizerVars::ox/oy
    bool in_float, out_float;
    int in_bytes, out_bytes; // sizeof(Tin), sizeof(Tout)
    bool use_gamma;
    int build_mode;    // -1 = auto
};

// The host half of resizeImage(): must make the same decisions, in the same order and on
// the same evolving bank state, as avir.h:4709-4954.
inline ImagePlan plan_image(Designer& d, const CallDesc& c) {
    ImagePlan ip;
    double kx, ky, ox = c.ox, oy = c.oy;
    if (c.k == 0.0) {
        kx = static_cast<double>(c.src_w) / c.new_w;
        ox += (kx - 1.0) * 0.5;
        ky = static_cast<double>(c.src_h) / c.new_h;
        oy += (ky - 1.0) * 0.5;
    } else if (c.k > 0.0) {
        kx = c.k; ky = c.k;
        const double ko = (c.k - 1.0) * 0.5;
        ox += ko; oy += ko;
    } else {
        kx = -c.k; ky = -c.k;
    }
    double out_mul;
    if (c.use_gamma) {
        ip.in_gamma_mult = c.in_float ? 1.0 : 1.0 / (c.in_bytes == 1 ? 255.0 : 65535.0);
        ip.out_gamma_mult = c.out_float ? 1.0 : (c.out_bytes == 1 ? 255.0 : 65535.0);
        out_mul = 1.0;
    } else {
        out_mul = c.out_float ? 1.0 : (c.out_bytes == 1 ? 255.0 : 65535.0);
        if (!c.in_float) out_mul /= (c.in_bytes == 1 ? 255.0 : 65535.0);
    }
    ip.kx = kx; ip.ky = ky; ip.ox = ox; ip.oy = oy; ip.out_mul = out_mul;
    const Mirror& m = d.mirror();
    const int el_count = (c.channels + m.fppack - 1) / m.fppack;
    ip.el_count = el_count;
    const int mode_count = d.fixed_order() == 0 ? 4 : 2;

    FracBank bank;
    std::vector<std::unique_ptr<Step>> steps;
    AxisVars v;

    // ---- horizontal axis
    int use_mode = 1;
    if (c.build_mode >= 0) {
        use_mode = c.build_mode;
    } else {
        int best = 0x7FFFFFFF;
        for (int mode = 0; mode < mode_count; ++mode) {
            FracBank tmp;
            std::vector<std::unique_ptr<Step>> ts;
            AxisVars tv;
            tv.k = kx; tv.o = ox;
            d.build_steps(ts, tv, tmp, out_mul, mode, true);
            d.size_steps(ts, tv, c.src_w, c.new_w);
            const int cost = d.complexity(ts, tv, el_count, c.src_h);
            if (cost < best) { use_mode = mode; best = cost; }
        }
    }
    v.k = kx; v.o = ox;
    d.build_steps(steps, v, bank, out_mul, use_mode, false);
    d.size_steps(steps, v, c.src_w, c.new_w);
    Designer::touch_phases(*steps[v.resize_step]);
    ip.h = flatten(steps, v, use_mode, c.src_w, c.new_w);

    // ---- vertical axis
    const int prev_mode = use_mode;
    if (c.build_mode >= 0) {
        use_mode = c.build_mode;
    } else {
        int best = 0x7FFFFFFF;
        for (int mode = 0; mode < mode_count; ++mode) {
            FracBank tmp;
            tmp.clone_params(bank);
            std::vector<std::unique_ptr<Step>> ts;
            AxisVars tv;
            tv.k = ky; tv.o = oy;
            d.build_steps(ts, tv, tmp, 1.0, mode, true);
            d.size_steps(ts, tv, c.src_h, c.new_h);
            const int cost = d.complexity(ts, tv, el_count, c.new_w);
            if (cost < best) { use_mode = mode; best = cost; }
        }
    }
    v.k = ky; v.o = oy;
    if (use_mode == prev_mode && ky == kx) {
        if (out_mul != 1.0) Designer::scale_correction(steps, 1.0 / out_mul);
    } else {
        d.build_steps(steps, v, bank, 1.0, use_mode, false);
    }
    d.size_steps(steps, v, c.src_h, c.new_h);
    Designer::touch_phases(*steps[v.resize_step]);
    ip.v = flatten(steps, v, use_mode, c.src_h, c.new_h);
    return ip;
}

} // namespace plan
} // namespace avirb200

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC pop_options
#endif

#endif // AVIRB200_PLAN_HPP
