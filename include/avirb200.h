/* avirb200.h -- C ABI of libavirb200.so, the B200 (sm_100a) execution engine behind the
 * header-only avir::CImageResizer<> / avir::CLancIR drop-in front-ends (avir_b200.h,
 * lancir_b200.h).
 *
 * Upstream AVIR has no FFI of its own (it is a header-only C++ template library), so this
 * boundary sits where a GPU can replace it: at the granularity of one
 * CImageResizer<>::resizeImage() call (upstream avir.h:4680-5092) resp. one
 * CLancIR::resizeImage() call (upstream lancir.h:386-713).  The host front-end does what
 * upstream does on the host -- decide the chain of filtering steps and design their
 * coefficients in double precision (upstream avir.h:5128-6270) -- and hands the result to
 * this library as a flat, pointer-and-size "plan descriptor".  The library owns only
 * device-side work: upload of the tables, the row pass, the column pass, the output
 * epilogue, and the inter-GPU halo exchange.
 *
 * Conventions: every function returns 0 on success or a negative avirb200_status; no
 * function throws; no function falls back to the CPU.  Plans are immutable after creation
 * and may be shared by threads/streams; a workspace belongs to one in-flight call.
 */
#ifndef AVIRB200_H
#define AVIRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum avirb200_status {
    AVIRB200_OK = 0,
    AVIRB200_ERR_BAD_ARG = -1,
    AVIRB200_ERR_CUDA = -2,
    AVIRB200_ERR_NCCL = -3,
    AVIRB200_ERR_UNSUPPORTED = -4,
    AVIRB200_ERR_NO_DEVICE = -5,
    AVIRB200_ERR_ALLOC = -6
} avirb200_status;

/* Element types of the caller's image buffers (upstream Tin/Tout, avir.h:4670-4677).
 * F64: upstream narrows double input with a (float) cast in packScanline and widens the float
 * result with a (double) cast in unpackScanline (avir.h:2803-2806, 3168-3171); the library
 * does the same casts on the device (resize_device / resize_host; the sharded calls and the
 * per-pass entry points take U8 / U16 / F32 only). */
typedef enum avirb200_dtype { AVIRB200_U8 = 0, AVIRB200_U16 = 1, AVIRB200_F32 = 2, AVIRB200_F64 = 3 } avirb200_dtype;

/* Step kinds of a 1-D filtering chain (upstream CImageResizerFilterStep, avir.h:2568-2728). */
typedef enum avirb200_step_kind {
    AVIRB200_STEP_FIR = 0,      /* doFilter  (avir.h:3748-3866, avir_dil.h:444-539) */
    AVIRB200_STEP_UPSAMPLE = 1, /* doUpsample with filtering (avir.h:3404-3733) */
    AVIRB200_STEP_RESIZE = 2    /* doResize / doResize2 (avir.h:3884-4328, avir_dil.h:559-761);
                                   a filterless 2X upsample in front of it is folded in */
} avirb200_step_kind;

/* Tap-accumulation order to mirror (this is what makes results bit-identical):
 *   INL  : upstream interleaved classes (fpclass_def<float>, fpclass_float4) -- sequential.
 *   DIL8 : upstream fpclass_float8_dil -- 8 lane-strided partial sums + float8::hadd tree
 *          (avir_dil.h:463-472,598-609; avir_float8_avx.h:264-273). */
typedef enum avirb200_sum_mode { AVIRB200_SUM_INL = 0, AVIRB200_SUM_DIL8 = 1 } avirb200_sum_mode;

/* Integer-output rounding to mirror (upstream CDitherer::dither + round()):
 *   HALFUP_INT : avir::round<float>, (int)(v+0.5) (avir.h:130-135)      -- fpclass_def<float>
 *   RNE_I32    : round(float4) through cvtps_epi32 (avir_float4_sse.h:303-313)
 *   RNE        : round(float8), _mm256_round_ps nearest-even (avir_float8_avx.h:347-351) */
typedef enum avirb200_round_mode {
    AVIRB200_ROUND_HALFUP_INT = 0,
    AVIRB200_ROUND_RNE_I32 = 1,
    AVIRB200_ROUND_RNE = 2
} avirb200_round_mode;

#define AVIRB200_MAX_STEPS 6

/* One filtering step.  All pointers are HOST pointers, copied by avirb200_plan_create. */
typedef struct avirb200_step_desc {
    int32_t kind;       /* avirb200_step_kind */
    int32_t resample;   /* FIR: decimation R >= 1; UPSAMPLE: factor (2) */
    int32_t latency;    /* FIR / UPSAMPLE: filter latency L */
    int32_t edge;       /* FIR: extra edge outputs per side (EdgePixelCount) */
    int32_t in_len;     /* input line length n; reads are clamped to [0, n-1] */
    int32_t out_len;    /* outputs produced */
    int32_t ntaps;      /* FIR/UPSAMPLE: taps stored; RESIZE: bank filter length FL */
    int32_t order;      /* RESIZE: 0 or 1 (c0 + c1*x) */
    int32_t upsampled;  /* RESIZE: input is the virtual 2X zero-stuffed line */
    int32_t skip_odd;   /* RESIZE: accumulate only taps that land on real samples (doResize2) */
    int32_t zero_start; /* accumulators start at +0 (avir.h:3938-3951, avir_dil.h:645) */
    int32_t nphases;    /* RESIZE: phases stored in taps[] */
    int32_t out_prefix; /* UPSAMPLE: outputs produced before position 0 */
    int32_t out_suffix; /* UPSAMPLE: outputs produced after position out_len-1 */
    int32_t in_prefix;  /* UPSAMPLE: times the first sample is additionally filtered */
    int32_t in_suffix;  /* UPSAMPLE: times the last sample is additionally filtered */
    int32_t n_prefix_dc, n_suffix_dc;
    const float* taps;       /* FIR/UPSAMPLE: ntaps; RESIZE: nphases*ntaps*(order+1) */
    const int32_t* src_pos;  /* RESIZE: out_len integer source positions (SrcPosInt) */
    const int32_t* phase;    /* RESIZE: out_len indices into taps[] phases */
    const float* frac;       /* RESIZE: out_len interpolation fractions x */
    const float* prefix_dc;  /* UPSAMPLE */
    const float* suffix_dc;  /* UPSAMPLE */
} avirb200_step_desc;

typedef struct avirb200_axis_desc {
    int32_t src_len, dst_len, nsteps;
    avirb200_step_desc steps[AVIRB200_MAX_STEPS];
} avirb200_axis_desc;

typedef struct avirb200_plan_desc {
    int32_t src_w, src_h, dst_w, dst_h;
    int32_t channels;        /* 1..4 (ElCountIO) */
    int32_t in_type, out_type; /* avirb200_dtype */
    int32_t sum_mode;        /* avirb200_sum_mode */
    int32_t round_mode;      /* avirb200_round_mode (integer output only) */
    int32_t use_gamma;       /* bit 0: sRGB-linearise the input; bit 1: de-linearise the output */
    int32_t alpha_index;     /* channel exempt from gamma when channels == 4 (0 or 3), else -1 */
    float in_gamma_mult;     /* (float)Vars.InGammaMult  (avir.h:2843) */
    float out_gamma_mult;    /* (float)Vars.OutGammaMult (avir.h:2987) */
    float tr_mul, tr_mul_inv; /* bit-depth truncation multipliers (avir.h:4408-4417); 1 = off */
    float pk_out;            /* output clamp ceiling (avir.h:5043) */
    int32_t dither;          /* integer output: 0 = per-sample rounding (CImageResizerDithererDefINL/DIL,
                                avir.h:4392-4419), 1 = error diffusion (CImageResizerDithererErrdINL/DIL,
                                avir.h:4442-4530, avir_dil.h:882-986): row-recursive, whole image only */
    avirb200_axis_desc h, v; /* row pass, column pass */
} avirb200_plan_desc;

typedef struct avirb200_plan avirb200_plan; /* opaque, device-resident tables */

/* ---- single-GPU path --------------------------------------------------------------- */

/* Copies the descriptor's tables to the current CUDA device. */
int avirb200_plan_create(const avirb200_plan_desc* desc, avirb200_plan** out);
void avirb200_plan_destroy(avirb200_plan* plan);

/* Bytes of device scratch one call needs (the fp32 row-pass intermediate,
 * upstream FltBuf, avir.h:4881-4883). */
int avirb200_plan_workspace_bytes(const avirb200_plan* plan, size_t* bytes);

/* Device-resident resize: src/dst/workspace are device pointers, pitches are in ELEMENTS
 * (upstream SrcScanlineSize semantics, avir.h:4647-4649).  Asynchronous on `stream`
 * (a cudaStream_t passed as void*); no allocation, no synchronisation. */
int avirb200_resize_device(const avirb200_plan* plan, const void* d_src, size_t src_pitch,
                           void* d_dst, size_t dst_pitch, void* d_workspace, void* stream);

/* The two passes of avirb200_resize_device individually (same arguments, same stream
 * semantics): row pass src -> workspace, column pass workspace -> dst.  For callers that
 * pipeline frames, and for per-kernel timing. */
int avirb200_row_pass_device(const avirb200_plan* plan, const void* d_src, size_t src_pitch,
                             void* d_workspace, void* stream);
int avirb200_col_pass_device(const avirb200_plan* plan, const void* d_workspace, void* d_dst,
                             size_t dst_pitch, void* stream);

/* Convenience used by the drop-in resizeImage(): host buffers in, host buffers out
 * (H2D, both passes, D2H, synchronise).  Device buffers are cached inside the plan.
 * Images of 64 MiB and more are cut into row bands travelling on separate copy-in / compute /
 * copy-out streams, so that the PCIe transfers of both directions overlap each other and the
 * kernels (page-locked host memory lets them run asynchronously); the result bits do not
 * depend on the banding.  h_dst may alias h_src (upstream allows NewBuf == SrcBuf when the
 * destination is not larger, avir.h:4650-4652): such calls run unbanded. */
int avirb200_resize_host(avirb200_plan* plan, const void* h_src, size_t src_pitch, void* h_dst,
                         size_t dst_pitch);

/* Number of kernel launches the last avirb200_resize_device on this plan issued. */
int avirb200_plan_last_launches(const avirb200_plan* plan);

/* Video / batch entry: `n` frames of the plan's geometry, one launch pair per frame, all on
 * `stream` in order (the frames share `d_workspace`: stream order keeps them apart).  d_srcs /
 * d_dsts are HOST arrays of n device pointers; pitches as in avirb200_resize_device.  The
 * plan-building cost (host filter design, table upload) is paid once for the whole batch --
 * upstream's equivalent is re-using one CImageResizer object and its Vars across frames. */
int avirb200_resize_device_batch(const avirb200_plan* plan, int n, const void* const* d_srcs,
                                 size_t src_pitch, void* const* d_dsts, size_t dst_pitch,
                                 void* d_workspace, void* stream);

/* Per-plan options (tuning and test switches; none changes a result bit).  Options are plan
 * state: set them while no call on the plan is in flight.  value < 0 restores the default. */
typedef enum avirb200_option {
    /* kernel family order: 0 = streaming kernel where the chain is regular, else the tile kernel,
     * else the generic kernel (default); 1 = generic kernel only; 2 = tile kernel, else generic */
    AVIRB200_OPT_KERNEL_FAMILY = 0,
    /* scheduling variant of the streaming row / column pass: 0 ring windows, 1 register windows,
     * 2 register windows + TMA-staged source ring (column pass) */
    AVIRB200_OPT_STREAM_VARIANT_H = 1,
    AVIRB200_OPT_STREAM_VARIANT_V = 2,
    /* avirb200_resize_host: number of row bands of the pipelined form (1 = unbanded; default by size) */
    AVIRB200_OPT_HOST_BANDS = 3,
    /* 1: also select the streaming chains that measured slower than the tile kernel (upsizing, 56-tap) */
    AVIRB200_OPT_ALL_STREAM_CHAINS = 4,
    /* avirb200_resize_sharded, how the halo rows travel.  3 (default) = the FUSED exchange: the row
     * kernel itself stores the rows the neighbours need into their mailboxes (peer stores over NVLink)
     * and raises their flags, the column kernel reads the neighbours' rows in place from the mailbox --
     * no exchange stream, no copies, no extra launch (a pass that is not on the streaming kernel falls
     * back to the push / pull of 1, per pass).  1 = rows pushed by the copy engines into the
     * neighbours' mailboxes after the row pass, pulled into the workspace by a small kernel; 2 = the
     * same with the rows the neighbours need filtered FIRST (one segmented launch) so that the push
     * overlaps the interior rows; 0 = NCCL send/recv between the two passes.  Measured, cfg3 weak
     * scaling on 2 x B200 (profiles/r02i_*): 0.2507 ms (3), 0.2609 ms (1); one GPU 0.2440 ms. */
    AVIRB200_OPT_OVERLAP_HALO = 5
} avirb200_option;
int avirb200_plan_set_option(avirb200_plan* plan, int option, int value);

/* ---- row-sharded multi-GPU path (one process per GPU) -------------------------------- */

/* Source/destination row ranges rank `rank` of `nranks` owns, and the intermediate rows it
 * must receive from its neighbours before the column pass. */
typedef struct avirb200_shard_info {
    int32_t src_row0, src_rows; /* source rows this rank filters in the row pass */
    int32_t dst_row0, dst_rows; /* destination rows this rank produces */
    int32_t need_row0, need_rows; /* intermediate rows the column pass reads */
    int32_t halo_up, halo_down;   /* rows received from rank-1 / rank+1 */
} avirb200_shard_info;

int avirb200_shard_query(const avirb200_plan* plan, int rank, int nranks,
                         avirb200_shard_info* info);
/* Same, from a descriptor alone (pure host arithmetic, no device needed). */
int avirb200_shard_query_desc(const avirb200_plan_desc* desc, int rank, int nranks,
                              avirb200_shard_info* info);
int avirb200_shard_workspace_bytes(const avirb200_plan* plan, int rank, int nranks,
                                   size_t* bytes);

/* NCCL bootstrap without exposing NCCL types: rank 0 fills a 128-byte id, the caller
 * broadcasts it by any means, every rank then creates its communicator. */
int avirb200_comm_unique_id(void* id128);
int avirb200_comm_create(const void* id128, int rank, int nranks, void** comm_out);
void avirb200_comm_destroy(void* comm);

/* d_src holds this rank's source band (src_rows rows), d_dst receives its destination band
 * (dst_rows rows).  Row pass -> NCCL halo send/recv with rank-1/rank+1 -> column pass, all
 * enqueued on `stream`.  `comm` is an ncclComm_t (from avirb200_comm_create or the
 * caller's own).  Output is bit-identical to the single-GPU path.
 * Default schedule (AVIRB200_OPT_OVERLAP_HALO = 3): every rank owns a mailbox in device memory that its
 * neighbours map through CUDA IPC (the handles travel over `comm` once per plan).  The row kernel
 * stores the rows a neighbour needs into that neighbour's mailbox as it produces them and, when the
 * last one is out, the call's sequence number into the neighbour's flag; the column kernel reads
 * the neighbours' rows in place from its own mailbox, and only the runs that touch them wait for
 * the flag.  The first call on a plan is collective (every rank must make it).  Where peer mapping
 * is unavailable the NCCL send/recv schedule runs. */
int avirb200_resize_sharded(const avirb200_plan* plan, void* comm, int rank, int nranks,
                            const void* d_src, size_t src_pitch, void* d_dst, size_t dst_pitch,
                            void* d_workspace, void* stream);

/* The same with HOST buffers (this rank's source band in, its destination band out): copies
 * in, avirb200_resize_sharded on the plan's own stream, copies out, synchronises.  Staging
 * buffers are cached in the plan.  The multi-GPU form of avirb200_resize_host. */
int avirb200_resize_sharded_host(avirb200_plan* plan, void* comm, int rank, int nranks,
                                 const void* h_src, size_t src_pitch, void* h_dst, size_t dst_pitch);

/* Validation aid: runs the `nranks` bands of the sharded schedule one after another on the
 * CURRENT device.  With the default AVIRB200_OPT_OVERLAP_HALO (3) the bands exchange their halo rows
 * exactly as ranks do -- the row kernel stores them into the neighbour band's mailbox (here in local
 * memory) and raises its flag, the column kernel reads them in place; otherwise with device copies.
 * Full-image device buffers; d_workspace must hold the sum of all ranks'
 * avirb200_shard_workspace_bytes. */
int avirb200_resize_sharded_local(const avirb200_plan* plan, int nranks, const void* d_src,
                                  size_t src_pitch, void* d_dst, size_t dst_pitch,
                                  void* d_workspace, void* stream);

/* Which specialised kernels the plan's passes qualify for: bit 0 / 1 = row / column pass on
 * the warp-streaming kernel, bit 2 / 3 = row / column pass on the tile kernel. */
int avirb200_plan_kernel_paths(const avirb200_plan* plan);

/* ---- LANCIR (upstream lancir.h) ------------------------------------------------------- */

typedef struct lancirb200_axis_desc {
    int32_t src_len, dst_len;
    int32_t kernel_len;     /* KernelLen (lancir.h:889-895) */
    int32_t nphases;
    const float* taps;      /* nphases * kernel_len, un-replicated (lancir.h:1076-1156) */
    const int32_t* src_pos; /* dst_len: first tap's source index (may be < 0 / >= src_len) */
    const int32_t* phase;   /* dst_len */
} lancirb200_axis_desc;

typedef struct lancirb200_plan_desc {
    int32_t src_w, src_h, dst_w, dst_h, channels;
    int32_t in_type, out_type;
    float out_mul;          /* lancir.h:526-533 */
    int32_t is_unity_mul;
    float clamp_max;        /* 255 / 65535 for integer output */
    lancirb200_axis_desc v, h; /* LANCIR resizes columns first, then rows */
} lancirb200_plan_desc;

typedef struct lancirb200_plan lancirb200_plan;

int lancirb200_plan_create(const lancirb200_plan_desc* desc, lancirb200_plan** out);
void lancirb200_plan_destroy(lancirb200_plan* plan);
int lancirb200_plan_workspace_bytes(const lancirb200_plan* plan, size_t* bytes);
int lancirb200_resize_device(const lancirb200_plan* plan, const void* d_src, size_t src_pitch,
                             void* d_dst, size_t dst_pitch, void* d_workspace, void* stream);
int lancirb200_resize_host(lancirb200_plan* plan, const void* h_src, size_t src_pitch,
                           void* h_dst, size_t dst_pitch);

/* ---- misc ----------------------------------------------------------------------------- */

const char* avirb200_status_string(int status);
const char* avirb200_last_error(void); /* thread-local detail of the last failure */
int avirb200_device_count(void);

/* Self-test on the current device: the batched, branch-free output gamma of the streaming column pass
 * (pixel_ops.cuh, lin2srgb_batch: the library square root's fast path spelled out so that several
 * samples' chains interleave) against the one-sample path (upstream avir.h:288-310 evaluated with
 * sqrt.rn.f64) on EVERY float bit pattern the batched path accepts.  *checked = patterns compared,
 * *mismatches = patterns whose results differ in any bit (must be 0). */
int avirb200_selftest_lin2srgb(unsigned long long* checked, unsigned long long* mismatches);

#ifdef __cplusplus
}
#endif

#endif /* AVIRB200_H */
