/* oracle/avir_port.c -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement ("port") of the
 * scanline arithmetic of upstream AVIR in gather form: every output sample is computed
 * independently from indexed reads, in exactly upstream's operation order.
 *
 * It executes the same flat plan descriptor (include/avirb200.h) the CUDA library executes,
 * so a mismatch between the two isolates a kernel bug, while the pinned oracle for the
 * whole pipeline (planner + arithmetic) remains oracle/_ref (upstream itself).  This port
 * is pinned by tests/test_oracle_port.py against oracle/_ref outputs and against the
 * committed golden fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's
 * shared object; the product never does.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction -- it changes the bits).
 *
 * Upstream code each function follows is cited as avir.h:<lines> / avir_dil.h:<lines>.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/avirb200.h"

typedef struct line {
    float* p; /* p[n] valid for lo <= n < hi */
    int lo, hi;
} line;

static inline float X(const line* l, int n)
{
    if (n < l->lo) n = l->lo;
    if (n >= l->hi) n = l->hi - 1;
    return l->p[n];
}

/* float8::hadd, avir_float8_avx.h:264-273 */
static inline float hadd8(const float* v)
{
    return ((v[0] + v[4]) + (v[1] + v[5])) + ((v[2] + v[6]) + (v[3] + v[7]));
}

/* doFilter: avir.h:3748-3866 (INL, folded symmetric form) */
static void fir_inl(const avirb200_step_desc* s, const line* in, float* out)
{
    const int L = s->latency;
    const float* f = s->taps + L;
    for (int j = 0; j < s->out_len; j++) {
        const int p = (j - s->edge) * s->resample;
        float sum = f[0] * X(in, p);
        for (int i = 1; i <= L; i++)
            sum += f[i] * (X(in, p + i) + X(in, p - i));
        out[j] = sum;
    }
}

/* doFilter: avir_dil.h:444-539 (DIL, full padded filter, lane-strided sums + hadd) */
static void fir_dil(const avirb200_step_desc* s, const line* in, float* out)
{
    for (int j = 0; j < s->out_len; j++) {
        const int p = (j - s->edge) * s->resample - s->latency;
        float lane[8];
        for (int q = 0; q < 8; q++)
            lane[q] = s->taps[q] * X(in, p + q);
        for (int i = 8; i < s->ntaps; i += 8)
            for (int q = 0; q < 8; q++)
                lane[q] += s->taps[i + q] * X(in, p + i + q);
        out[j] = hadd8(lane);
    }
}

/* Sample of the (possibly virtual 2X zero-stuffed) input line of a resize step.
 * Filterless doUpsample: avir.h:3260-3402, avir_dil.h:322-358. */
static inline float XR(const avirb200_step_desc* s, const line* in, int n)
{
    if (!s->upsampled)
        return X(in, n);
    if (n & 1)
        return 0.0f;
    return X(in, n >> 1); /* arithmetic shift == floor for negative n */
}

/* doResize / doResize2: avir.h:3884-4328 (INL) */
static void resize_inl(const avirb200_step_desc* s, const line* in, float* out)
{
    const int FL = s->ntaps;
    for (int j = 0; j < s->out_len; j++) {
        const float* c0 = s->taps + (size_t)s->phase[j] * FL * (s->order + 1);
        const float* c1 = c0 + FL;
        const float x = s->frac[j];
        const int p = s->src_pos[j] - (FL / 2 - 1);
        float sum = 0.0f;
        int first = !s->zero_start;
        for (int i = 0; i < FL; i++) {
            if (s->skip_odd && ((p + i) & 1))
                continue;
            const float t = s->order ? c0[i] + c1[i] * x : c0[i];
            const float v = t * XR(s, in, p + i);
            if (first) { sum = v; first = 0; }
            else sum += v;
        }
        out[j] = sum;
    }
}

/* doResize: avir_dil.h:559-751 (DIL) */
static void resize_dil(const avirb200_step_desc* s, const line* in, float* out)
{
    const int FL = s->ntaps;
    for (int j = 0; j < s->out_len; j++) {
        const float* c0 = s->taps + (size_t)s->phase[j] * FL * (s->order + 1);
        const float* c1 = c0 + FL;
        const float x = s->frac[j];
        const int p = s->src_pos[j] - (FL / 2 - 1);
        float lane[8];
        for (int q = 0; q < 8; q++) lane[q] = 0.0f;
        for (int i = 0; i < FL; i += 8) {
            for (int q = 0; q < 8; q++) {
                const float t = s->order ? c0[i + q] + c1[i + q] * x : c0[i + q];
                const float v = t * XR(s, in, p + i + q);
                if (i == 0 && !s->zero_start) lane[q] = v;
                else lane[q] += v;
            }
        }
        out[j] = hadd8(lane);
    }
}

static inline int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

/* doUpsample with filtering: avir.h:3404-3733, avir_dil.h:360-429.  Upstream scatters
 * taps of every (edge-replicated) input sample into a zeroed line; in gather form the
 * contributions to one output are added in order of increasing input index, then the
 * suffix tail, then the prefix tail.  `out` is indexed from -out_prefix. */
static void upsample_filtered(const avirb200_step_desc* s, const line* in, float* out)
{
    const int R = s->resample;
    const int first_m = -s->in_prefix;
    const int last_m = s->in_len - 1 + s->in_suffix;
    const int sfx_base = (last_m + 1) * R - s->latency; /* where the suffix tail lands */
    const int pfx_base = -s->in_prefix * R;
    for (int n = -s->out_prefix; n < s->out_len + s->out_suffix; n++) {
        float sum = 0.0f;
        /* taps index i = n - m*R + latency must lie in [0, ntaps) */
        int m_lo = fdiv(n + s->latency - (s->ntaps - 1) + R - 1, R);
        int m_hi = fdiv(n + s->latency, R);
        if (m_lo < first_m) m_lo = first_m;
        if (m_hi > last_m) m_hi = last_m;
        for (int m = m_lo; m <= m_hi; m++)
            sum += s->taps[n - m * R + s->latency] * X(in, m);
        if (n >= sfx_base && n < sfx_base + s->n_suffix_dc)
            sum += X(in, s->in_len - 1) * s->suffix_dc[n - sfx_base];
        if (n >= pfx_base && n < pfx_base + s->n_prefix_dc)
            sum += X(in, 0) * s->prefix_dc[n - pfx_base];
        out[n] = sum;
    }
}

/* Runs the whole chain of one axis on one single-channel line.
 * resizeScanlineH/V: avir.h:6522-6619. */
static void run_chain(const avirb200_axis_desc* ax, int sum_mode, const float* src, float* dst,
                      float* bufA, float* bufB, int guard)
{
    line cur;
    cur.p = (float*)src;
    cur.lo = 0;
    cur.hi = ax->src_len;
    float* bufs[2] = { bufA, bufB };
    for (int i = 0; i < ax->nsteps; i++) {
        const avirb200_step_desc* s = &ax->steps[i];
        const int last = (i == ax->nsteps - 1);
        float* o = last ? dst : bufs[i & 1] + guard;
        line nxt;
        nxt.p = o;
        nxt.lo = 0;
        nxt.hi = s->out_len;
        switch (s->kind) {
        case AVIRB200_STEP_FIR:
            if (sum_mode == AVIRB200_SUM_DIL8) fir_dil(s, &cur, o);
            else fir_inl(s, &cur, o);
            break;
        case AVIRB200_STEP_RESIZE:
            if (sum_mode == AVIRB200_SUM_DIL8) resize_dil(s, &cur, o);
            else resize_inl(s, &cur, o);
            break;
        case AVIRB200_STEP_UPSAMPLE:
            upsample_filtered(s, &cur, o);
            nxt.lo = -s->out_prefix;
            nxt.hi = s->out_len + s->out_suffix;
            break;
        }
        cur = nxt;
    }
}

/* ---- sRGB helpers: avir.h:162-310 (double-precision polynomials, float I/O) ------------- */

static float pow24_srgb(float x0)
{
    const double x = (double)x0;
    const double x2 = x * x;
    const double x3 = x2 * x;
    const double x4 = x2 * x2;
    return (float)(0.0985766365536824 + 0.839474952656502 * x2 + 0.363287814061725 * x3 -
                   0.0125559718896615 / (0.12758338921578 + 0.290283465468235 * x) -
                   0.231757513261358 * x - 0.0395365717969074 * x4);
}

static float pow24i_srgb(float x0)
{
    const double x = (double)x0;
    const double sx = sqrt(x);
    const double ssx = sqrt(sx);
    const double sssx = sqrt(ssx);
    return (float)(0.000213364515060263 + 0.0149409239419218 * x + 0.433973412731747 * sx +
                   ssx * (0.659628181609715 * sssx - 0.0380957908841466 -
                          0.0706476137208521 * sx));
}

static float srgb2lin_f(float s0, float m)
{
    const float s = s0 * m;
    const float a = 0.055f;
    if (s <= 0.04045f)
        return s / 12.92f;
    return pow24_srgb((s + a) / (1.0f + a));
}

/* The u8 table upstream ships (avir.h:234-286) equals the double-precision formula
 * rounded to 7 significant decimal digits, then parsed as a float literal. */
static float srgb2lin_u8(int v)
{
    const double sv = v / 255.0;
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
    /* %.7g in the C locale */
    int n = 0;
    {
        extern int snprintf(char*, size_t, const char*, ...);
        n = snprintf(buf, sizeof buf, "%.7g", r);
    }
    (void)n;
    return strtof(buf, NULL);
}

static float lin2srgb(float s)
{
    const float a = 0.055f;
    if (s <= 0.0031308f)
        return 12.92f * s;
    return (1.0f + a) * pow24i_srgb(s) - a;
}

/* round(): avir.h:130-135, avir_float4_sse.h:303-313, avir_float8_avx.h:347-351 */
static float round_mode(float v, int mode)
{
    if (mode == AVIRB200_ROUND_HALFUP_INT)
        return v < 0.0f ? -(float)(int)(0.5f - v) : (float)(int)(v + 0.5f);
    if (mode == AVIRB200_ROUND_RNE_I32) {
        if (!(v >= -2147483648.0f && v < 2147483648.0f))
            return -2147483648.0f; /* cvtps_epi32 "integer indefinite" */
        return (float)(int)nearbyintf(v);
    }
    return nearbyintf(v);
}

static float load_elem(const void* src, int type, size_t idx)
{
    switch (type) {
    case AVIRB200_U8: return (float)((const uint8_t*)src)[idx];
    case AVIRB200_U16: return (float)((const uint16_t*)src)[idx];
    case AVIRB200_F64: return (float)((const double*)src)[idx]; /* (fptypeatom) ip[c], avir.h:2803-2806 */
    default: return ((const float*)src)[idx];
    }
}

typedef struct scratch {
    float *lin, *lout, *bufA, *bufB;
    int guard;
} scratch;

static int scratch_init(scratch* sc, const avirb200_plan_desc* d)
{
    int maxlen = d->src_w > d->src_h ? d->src_w : d->src_h;
    sc->guard = 0;
    for (int a = 0; a < 2; a++) {
        const avirb200_axis_desc* ax = a ? &d->v : &d->h;
        for (int i = 0; i < ax->nsteps; i++) {
            const int l = ax->steps[i].out_len + ax->steps[i].out_prefix + ax->steps[i].out_suffix;
            if (l > maxlen) maxlen = l;
            if (ax->steps[i].out_prefix > sc->guard) sc->guard = ax->steps[i].out_prefix;
        }
    }
    sc->lin = (float*)malloc((size_t)(maxlen + 16) * sizeof(float));
    sc->lout = (float*)malloc((size_t)(maxlen + 16) * sizeof(float));
    sc->bufA = (float*)malloc((size_t)(maxlen + sc->guard + 16) * sizeof(float));
    sc->bufB = (float*)malloc((size_t)(maxlen + sc->guard + 16) * sizeof(float));
    return (sc->lin && sc->lout && sc->bufA && sc->bufB) ? 0 : AVIRB200_ERR_ALLOC;
}

static void scratch_free(scratch* sc)
{
    free(sc->lin); free(sc->lout); free(sc->bufA); free(sc->bufB);
}

/* Row pass over `rows` source rows starting at `src` (packScanline: avir.h:2777-2971,
 * avir_dil.h:64-115; then the H chain) into mid[rows][dst_w*C]. */
int avir_port_row_pass(const avirb200_plan_desc* d, const void* src, size_t src_pitch, int rows,
                       float* mid)
{
    const int C = d->channels, sw = d->src_w, dw = d->dst_w;
    scratch sc;
    if (scratch_init(&sc, d) != 0) return AVIRB200_ERR_ALLOC;
    float lut[256];
    if ((d->use_gamma & 1) && d->in_type == AVIRB200_U8)
        for (int i = 0; i < 256; i++) lut[i] = srgb2lin_u8(i);
    for (int y = 0; y < rows; y++) {
        for (int c = 0; c < C; c++) {
            for (int x = 0; x < sw; x++) {
                const size_t idx = (size_t)y * src_pitch + (size_t)x * C + c;
                float v;
                if (!(d->use_gamma & 1)) {
                    v = load_elem(src, d->in_type, idx);
                } else if (C == 4 && c == d->alpha_index) {
                    v = load_elem(src, d->in_type, idx) * d->in_gamma_mult;
                } else if (d->in_type == AVIRB200_U8) {
                    v = lut[((const uint8_t*)src)[idx]];
                } else {
                    v = srgb2lin_f(load_elem(src, d->in_type, idx), d->in_gamma_mult);
                }
                sc.lin[x] = v;
            }
            run_chain(&d->h, d->sum_mode, sc.lin, sc.lout, sc.bufA, sc.bufB, sc.guard);
            for (int x = 0; x < dw; x++)
                mid[((size_t)y * dw + x) * C + c] = sc.lout[x];
        }
    }
    scratch_free(&sc);
    return 0;
}

/* Error-diffusion dithering of the finished (gamma-corrected) float rows, top to bottom:
 * CImageResizerDithererErrdINL::dither (avir.h:4485-4525) / ErrdDIL (avir_dil.h:927-986),
 * driven row by row from resizeImage (avir.h:5046-5064).  Per channel, pixel j of a row:
 *   R = (v[j] + Dith[j]) [+ 0.364842 * Noise(j-1)];  z = round(R * TrMulI) * TrMul;
 *   Noise = R - z;  out = clamp(z, 0, PkOut);
 *   Dith'[j-1] += 0.207305 * Noise;  Dith'[j] += 0.364842 * Noise;  Dith'[j+1] += 0.063011 * Noise
 * where Dith' (zeroed when the row begins) is what the NEXT row adds.  Contributions reach a
 * Dith' element in the order (from j-1: 0.063011) -> (own: 0.364842) -> (from j+1: 0.207305),
 * starting from 0; the additions are kept in that order.  round() is the class's own.
 * The de-interleaved class keeps the channels of a row as consecutive planes and runs them
 * one after the other, so its "Dith'[j-1] +=" of pixel 0 of channel c+1 lands on the LAST
 * pixel of channel c (avir_dil.h:964: rsdj[-1] with j = 0): mirrored. */
static void errd_rows(const avirb200_plan_desc* d, float* res, void* dst, size_t dst_pitch)
{
    const int C = d->channels, W = d->dst_w, H = d->dst_h;
    const float c1 = 0.364842f, c2 = 0.207305f, c3 = 0.063011f;
    float* dith = (float*)calloc((size_t)(W + 2) * C, sizeof(float)); /* one pixel of slack either side */
    float* D = dith + C;
    const int planar = (d->sum_mode == AVIRB200_SUM_DIL8);
    for (int y = 0; y < H; y++) {
        float* r = res + (size_t)y * W * C;
        float first_n2[4] = {0, 0, 0, 0};
        int leak[4] = {0, 0, 0, 0};
        for (int j = 0; j < W * C; j++) { /* avir.h:4493-4497 */
            r[j] = r[j] + D[j];
            D[j] = 0.0f;
        }
        for (int j = 0; j < W * C; j++) {
            const float z0 = round_mode(r[j] * d->tr_mul_inv, d->round_mode) * d->tr_mul;
            const float noise = r[j] - z0;
            r[j] = z0 < 0.0f ? 0.0f : (z0 > d->pk_out ? d->pk_out : z0);
            if (j < C) first_n2[j] = noise * c2;
            if (j < (W - 1) * C) { /* avir.h:4499-4513 */
                const float nm1 = noise * c1;
                r[j + C] = r[j + C] + nm1;
                D[j - C] = D[j - C] + noise * c2;
                D[j] = D[j] + nm1;
                D[j + C] = D[j + C] + noise * c3;
            } else { /* the last pixel, avir.h:4515-4524 */
                D[j - C] = D[j - C] + noise * c2;
                D[j] = D[j] + noise * c1;
            }
            if (planar && j >= (W - 1) * C && j % C + 1 < C) /* leak from the next plane's pixel 0 */
                leak[j % C] = 1;
            const size_t idx = (size_t)y * dst_pitch + j;
            if (d->out_type == AVIRB200_U8) ((uint8_t*)dst)[idx] = (uint8_t)r[j];
            else ((uint16_t*)dst)[idx] = (uint16_t)r[j];
        }
        for (int c = 0; c + 1 < C; c++) /* after the whole row: plane c+1 ran after plane c */
            if (leak[c]) D[(size_t)(W - 1) * C + c] = D[(size_t)(W - 1) * C + c] + first_n2[c + 1];
    }
    free(dith);
}

/* Column pass + epilogue for destination rows [out0, out1).  `mid` holds intermediate rows
 * [mid_row0, mid_row0 + mid_rows) of the image; rows outside are poisoned with NaN, so a
 * band that does not contain everything the outputs depend on is detected (returns
 * the number of non-finite results, 0 = ok).  `dst` row 0 is destination row out0.
 * Epilogue: applySRGBGamma (avir.h:2982-3068) -> dither (avir.h:4392-4419,
 * avir_dil.h:815-859) -> unpackScanline (avir.h:3155-3215). */
int avir_port_col_pass(const avirb200_plan_desc* d, const float* mid, int mid_row0, int mid_rows,
                       int out0, int out1, void* dst, size_t dst_pitch)
{
    const int C = d->channels, sh = d->src_h, dw = d->dst_w;
    scratch sc;
    /* error diffusion (integer output): row-recursive from row 0, whole image only */
    const int errd = (d->dither == 1 && d->out_type != AVIRB200_F32 && d->out_type != AVIRB200_F64);
    float* res = NULL;
    if (errd) {
        if (out0 != 0 || out1 != d->dst_h) return AVIRB200_ERR_UNSUPPORTED;
        res = (float*)malloc((size_t)d->dst_h * dw * C * sizeof(float));
        if (!res) return AVIRB200_ERR_ALLOC;
    }
    if (scratch_init(&sc, d) != 0) { free(res); return AVIRB200_ERR_ALLOC; }
    int bad = 0;
    for (int x = 0; x < dw; x++) {
        for (int c = 0; c < C; c++) {
            for (int y = 0; y < sh; y++)
                sc.lin[y] = (y >= mid_row0 && y < mid_row0 + mid_rows)
                                ? mid[((size_t)(y - mid_row0) * dw + x) * C + c] : NAN;
            run_chain(&d->v, d->sum_mode, sc.lin, sc.lout, sc.bufA, sc.bufB, sc.guard);
            for (int y = out0; y < out1; y++) {
                float v = sc.lout[y];
                if (!isfinite(v)) bad++;
                if (d->use_gamma & 2) {
                    if (C == 4 && c == d->alpha_index) v = v * d->out_gamma_mult;
                    else v = lin2srgb(v) * d->out_gamma_mult;
                }
                const size_t idx = (size_t)(y - out0) * dst_pitch + (size_t)x * C + c;
                if (errd) {
                    res[((size_t)y * dw + x) * C + c] = v;
                    continue;
                }
                if (d->out_type == AVIRB200_F32) {
                    ((float*)dst)[idx] = v;
                    continue;
                }
                if (d->out_type == AVIRB200_F64) { /* (Tout) v[c], avir.h:3168-3171 */
                    ((double*)dst)[idx] = (double)v;
                    continue;
                }
                if (d->tr_mul == 1.0f) v = round_mode(v, d->round_mode);
                else v = round_mode(v * d->tr_mul_inv, d->round_mode) * d->tr_mul;
                v = v < 0.0f ? 0.0f : (v > d->pk_out ? d->pk_out : v);
                if (d->out_type == AVIRB200_U8) ((uint8_t*)dst)[idx] = (uint8_t)v;
                else ((uint16_t*)dst)[idx] = (uint16_t)v;
            }
        }
    }
    scratch_free(&sc);
    if (errd) {
        errd_rows(d, res, dst, dst_pitch);
        free(res);
    }
    return bad;
}

/* Whole image: row pass -> fp32 intermediate -> column pass.  resizeImage: avir.h:4680-5092. */
int avir_port_resize(const avirb200_plan_desc* d, const void* src, size_t src_pitch, void* dst,
                     size_t dst_pitch)
{
    float* mid = (float*)malloc((size_t)d->dst_w * d->src_h * d->channels * sizeof(float));
    if (!mid) return AVIRB200_ERR_ALLOC;
    int r = avir_port_row_pass(d, src, src_pitch, d->src_h, mid);
    if (r == 0) {
        r = avir_port_col_pass(d, mid, 0, d->src_h, 0, d->dst_h, dst, dst_pitch);
        if (r > 0) r = 0; /* non-finite data is the caller's business for whole images */
    }
    free(mid);
    return r;
}

/* The u8 sRGB table, for the host-logic tests. */
void avir_port_srgb_lut(float* out256)
{
    for (int i = 0; i < 256; i++) out256[i] = srgb2lin_u8(i);
}

/* ---- LANCIR (upstream lancir.h, AVX build), 1-4 channel restatement -----------------------
 * Column pass first (lancir.h:603-646), then row pass + output (lancir.h:650-706).  Both
 * passes run the same resizeN tap loop; its summation tree depends on the channel count:
 *   C=4 (resize4, lancir.h:2466-2515): even taps and odd taps in two chains, added at the end;
 *   C=1 (resize1, lancir.h:2101-2214) and C=2 (resize2, 2216-2320): four chains S0..S3 over
 *     taps t = j mod 4 of the first kl&~3 taps; when kl%4 == 2 the last two products Pa, Pb
 *     join as ((S0+S2)+Pa) + ((S1+S3)+Pb), else (S0+S2)+(S1+S3);
 *   C=3 (resize3, 2322-2440): same four chains; Pa joins chain 0 before the tree; channel 0
 *     also folds Pb into chain 1 (lane 3 of the 128-bit accumulator), channels 1 and 2 add
 *     Pb last: c0 = (S0'+S1')+(S2+S3), c1,c2 = ((S0'+S1)+(S2+S3))+Pb.
 * Output: lancir.h:1772-2056 (vector body rounds to nearest-even, the last (W*C)&3
 * elements of a row use (int)(v+0.5f)). */
static float lancir_tapsum(const float* f, int kl, const float* base, long long stride, int s0,
                           int n, int C, int c)
{
#define LTAP(t) (f[t] * base[(long long)((s0 + (t)) < 0 ? 0 : ((s0 + (t)) >= n ? n - 1 : (s0 + (t)))) * stride])
    if (C == 4) {
        float ev = LTAP(0), od = LTAP(1);
        for (int t = 2; t < kl; t += 2) {
            ev += LTAP(t);
            od += LTAP(t + 1);
        }
        return ev + od;
    }
    const int n4 = kl & ~3;
    float s[4];
    for (int j = 0; j < 4; j++) s[j] = LTAP(j);
    for (int t = 4; t < n4; t += 4)
        for (int j = 0; j < 4; j++) s[j] += LTAP(t + j);
    const int rem = (kl & 3) == 2;
    float pa = 0.0f, pb = 0.0f;
    if (rem) { pa = LTAP(n4); pb = LTAP(n4 + 1); }
#undef LTAP
    if (C == 3) {
        if (rem) s[0] += pa;
        if (c == 0) {
            if (rem) s[1] += pb;
            return (s[0] + s[1]) + (s[2] + s[3]);
        }
        float r = (s[0] + s[1]) + (s[2] + s[3]);
        if (rem) r += pb;
        return r;
    }
    float a = s[0] + s[2], b = s[1] + s[3];
    if (rem) { a += pa; b += pb; }
    return a + b;
}

int lancir_port_resize(const lancirb200_plan_desc* d, const void* src, size_t src_pitch, void* dst,
                       size_t dst_pitch)
{
    const int C = d->channels, sw = d->src_w, sh = d->src_h, dw = d->dst_w, dh = d->dst_h;
    float* in = (float*)malloc((size_t)sw * sh * C * sizeof(float));
    float* mid = (float*)malloc((size_t)sw * dh * C * sizeof(float));
    if (!in || !mid) return AVIRB200_ERR_ALLOC;
    for (int y = 0; y < sh; y++)
        for (int e = 0; e < sw * C; e++)
            in[(size_t)y * sw * C + e] = load_elem(src, d->in_type, (size_t)y * src_pitch + e);
    for (int y = 0; y < dh; y++) {
        const float* f = d->v.taps + (size_t)d->v.phase[y] * d->v.kernel_len;
        for (int e = 0; e < sw * C; e++)
            mid[(size_t)y * sw * C + e] =
                lancir_tapsum(f, d->v.kernel_len, in + e, (long long)sw * C, d->v.src_pos[y], sh, C,
                              e % C);
    }
    const int out_elems = dw * C;
    for (int y = 0; y < dh; y++) {
        for (int x = 0; x < dw; x++) {
            const float* f = d->h.taps + (size_t)d->h.phase[x] * d->h.kernel_len;
            for (int c = 0; c < C; c++) {
                float v = lancir_tapsum(f, d->h.kernel_len, mid + (size_t)y * sw * C + c, C,
                                        d->h.src_pos[x], sw, C, c);
                if (!d->is_unity_mul) v = v * d->out_mul;
                const int e = x * C + c;
                const size_t idx = (size_t)y * dst_pitch + e;
                if (d->out_type == AVIRB200_F32) { ((float*)dst)[idx] = v; continue; }
                int iv;
                if (e >= (out_elems & ~3)) {
                    const float cv = v > d->clamp_max ? d->clamp_max : (v < 0.0f ? 0.0f : v);
                    iv = (int)(cv + 0.5f);
                } else {
                    float cv = v < d->clamp_max ? v : d->clamp_max;
                    cv = cv > 0.0f ? cv : 0.0f;
                    iv = (int)nearbyintf(cv);
                }
                if (d->out_type == AVIRB200_U8) ((uint8_t*)dst)[idx] = (uint8_t)iv;
                else ((uint16_t*)dst)[idx] = (uint16_t)iv;
            }
        }
    }
    free(in); free(mid);
    return 0;
}
