/*
 * b200nn.h -- C-ABI of the B200 (sm_100a) operator shim behind the CSI-NN2 API.
 *
 * This is the device-side half of the drop-in boundary for the hot path named in
 * BASELINE.json: csinn_conv2d (im2col + GEMM), csinn_depthwise_conv2d,
 * csinn_fullyconnected and the bandwidth-bound ops around them.  Everything is
 * `extern "C"`, plain pointers and sizes: no csinn_* type and no torch type
 * crosses this line.  The C host side (csi-nn2_b200/b200_opt/ -- the code that
 * registers in the reference's backend registry, see include/shl_b200.h) is the
 * only caller in the product; tests call the same symbols through ctypes.
 *
 * Each entry point cites the reference interface it stands in for
 * (paths relative to the reference tree).
 *
 * Conventions
 *   - Device activations are "pixel-major": [N][H][W][Cp] with Cp = channel
 *     stride in elements, Cp*elem a multiple of 16 bytes (b200_round_channels()).
 *     The API-side NCHW tensors of the reference
 *     (include/csinn/csinn_data_structure.h:505) are converted at the boundary by
 *     b200_nchw_to_nhwc / b200_nhwc_to_nchw.
 *   - dtype: B200_I8 (CSINN_DTYPE_INT8) or B200_F16 (CSINN_DTYPE_FLOAT16).
 *   - All functions return 0 on success, a negative b200_status otherwise, and
 *     never fall back to the CPU.  b200_last_error() gives the message.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Launches are asynchronous on `stream`; descriptors are copied at launch.
 *
 * Quantised epilogue (the one arithmetic contract; oracle/oracle_int.c restates it):
 *     acc  = sum_k x~[k] * w[o][k]                        (int32, exact; x~ = zp_in at pads)
 *     acc += ibias[o]                                     (int32: -zp_in * sum_k w[o][k])
 *     acc -= w_zp[o] * rs                                 (asymmetric weights only: rs = sum_k x~[k] - zp_in * K,
 *                                                          so that acc = sum_k (x~ - zp_in) * (w - w_zp[o]);
 *                                                          source/nn2/utils.c:920-931 dequantises kernels so)
 *     f    = fmaf((float)acc, mult[o], badd[o])           (one rounding)
 *     q    = clamp((int)rintf(f) + zp_out, -128, 127)     (round-half-even, like nearbyint in
 *                                                          source/nn2/utils.c:550 float_to_int8_base)
 *     act  : B200_ACT_RELU  -> q = max(q, zp_out);  B200_ACT_RELU6 -> also min(q, q6)
 *     post : optional 256-entry table applied last, q = post_lut[q + 128].  It holds a
 *            standalone relu / relu6 node with its own qinfo evaluated exactly as
 *            source/reference/relu.c:39 does through utils.c:609 siso_callback_base:
 *              r = ((float)q - zp_in) * s_in ; r = act(r) ; q' = clamp(nearbyint(r / s_out) + zp_out)
 *            (b200_build_requant_lut builds it on the host).
 *   fp16:  f = acc_f32 + badd[o] ; act(f) ; __float2half_rn.
 */
#ifndef B200NN_H_
#define B200NN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NN_ABI_VERSION 3

typedef enum {
    B200_OK = 0,
    B200_ERR_CUDA = -1,        /* a CUDA runtime/driver call failed */
    B200_ERR_UNSUPPORTED = -2, /* shape / dtype outside what the kernels cover */
    B200_ERR_ARG = -3,         /* malformed descriptor */
    B200_ERR_NO_DEVICE = -4,   /* no sm_100 device visible: there is NO cpu fallback */
} b200_status;

typedef enum { B200_I8 = 0, B200_F16 = 1 } b200_dtype;
typedef enum {
    B200_ACT_NONE = 0,
    B200_ACT_RELU = 1,
    B200_ACT_RELU6 = 2,
    /* unary ops that exist only as tables (int8) or in the fp16 elementwise kernel -- never as an
     * in-epilogue clamp: source/reference/leaky_relu.c:33, sigmoid.c:33, clip.c:32-38 */
    B200_ACT_LEAKY_RELU = 3, /* p0 = negative slope */
    B200_ACT_SIGMOID = 4,
    B200_ACT_CLIP = 5,       /* p0 = min, p1 = max */
    B200_ACT_SILU = 6,       /* val / (1.0f + exp(-val)), source/reference/silu.c:31 */
    B200_ACT_ERF = 7         /* erf(val), source/reference/erf.c:31 */
} b200_act;

/* Requantisation / epilogue parameters shared by conv, depthwise, fc.
 * For B200_F16 only `badd` (bias as float, may be NULL) and `act` are used. */
typedef struct {
    const float *mult;      /* device [O]  s_in*s_w[o]/s_out                          */
    const float *badd;      /* device [O]  bias_q[o]*s_bias[o]/s_out (f16: bias)      */
    const int32_t *ibias;   /* device [O]  integer addend (zero-point fold), or NULL  */
    const int8_t *post_lut; /* device [256] or NULL                                   */
    int32_t zp_out;
    int32_t act; /* b200_act */
    int32_t q6;  /* quantised value of 6.0 in the output domain (RELU6)                */
} b200_epilogue;

/* ---- device / memory plumbing ------------------------------------------------ */
int b200_abi_version(void);
const char *b200_last_error(void);
int b200_device_count(void);
int b200_set_device(int dev);
int b200_sm_count(void);
int b200_malloc(void **dptr, size_t bytes);
int b200_free(void *dptr);
int b200_malloc_host(void **hptr, size_t bytes); /* pinned */
int b200_free_host(void *hptr);
int b200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int b200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
/* `rows` rows of row_bytes, src_pitch apart on the device, packed densely on the host: a pixel-major
 * [n][1][1][cp] tensor IS the NCHW tensor [n][c] up to the row padding, so such a graph output is
 * read back with the copy engine alone (no compaction kernel) */
int b200_memcpy_d2h_rows(void *dst, const void *src, size_t row_bytes, size_t src_pitch, size_t rows, void *stream);
int b200_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int b200_memset(void *dst, int value, size_t bytes, void *stream);
int b200_stream_create(void **stream);
int b200_stream_destroy(void *stream);
int b200_stream_sync(void *stream);
int b200_device_sync(void);
/* CUDA events on `stream` (bench.py times kernels on the launching stream) */
int b200_event_create(void **ev);
int b200_event_destroy(void *ev);
int b200_event_record(void *ev, void *stream);
int b200_event_sync(void *ev);
int b200_event_elapsed_ms(void *start, void *stop, float *ms);
/* make `stream` wait for an event recorded on another stream (copy/compute overlap) */
int b200_stream_wait_event(void *stream, void *ev);
/* CUDA-graph capture of a launch sequence (graph-mode session_run) */
/* programmatic dependent launch for the kernels launched from now on (process-wide switch; the
 * graph-mode session sets it around its captures, see csi-nn2_b200/csrc/runtime.cu) */
void b200_set_pdl(int on);
int b200_get_pdl(void);
int b200_graph_begin(void *stream);
int b200_graph_end(void *stream, void **graph_exec);
int b200_graph_launch(void *graph_exec, void *stream);
int b200_graph_destroy(void *graph_exec);
/* number of kernels this library has launched since load (bench.py "gpu_launches");
 * launches replayed through a CUDA graph are counted per replay */
uint64_t b200_launch_count(void);
/* evict L2: overwrite an internal >L2-sized scratch buffer (bench.py timing hygiene) */
int b200_flush_l2(void *stream);

static inline int b200_round_channels(int c, int elem_bytes)
{
    int per16 = 16 / elem_bytes;
    return (c + per16 - 1) / per16 * per16;
}

/* host helper: the 256-entry requantisation table described above (index = q + 128) */
/* the reference's dequant -> f32 unary op -> requant (shl_ref_siso_callback_base, source/reference/utils.c:609)
 * as a function of the int8 input, evaluated on the host with the same float / libm sequence */
void b200_build_unary_lut(int8_t lut[256], int act, float p0, float p1, float s_in, int zp_in, float s_out, int zp_out);
void b200_build_requant_lut(int8_t lut[256], int act, float s_in, int zp_in, float s_out,
                            int zp_out);

/* ---- layout conversion at the API boundary ----------------------------------- */
/* NCHW [N][C][H][W] -> pixel-major [N][H][W][Cp]; pad channels are written as `pad`.
 * Stands in for the NCHW<->NC1HWC0 reorders of the RVV back end
 * (source/thead_rvv/data_convert.c, source/c920_opt/setup.c:316-352). */
int b200_nchw_to_nhwc(const void *src, void *dst, int n, int c, int h, int w, int cp,
                      int elem_bytes, int pad, void *stream);
int b200_nhwc_to_nchw(const void *src, void *dst, int n, int c, int h, int w, int cp,
                      int elem_bytes, void *stream);

/* ---- GEMM on tcgen05: 1x1 conv2d, fullyconnected, and the GEMM half of im2col conv */
/* out[m][o] = epilogue( sum_k a[m][k] * w[o][k] ),  m = pixel (N*H*W) or fc batch row.
 * Replaces shl_rvv_conv1x1s1_gemm_int8 (source/thead_rvv/int8/convolution_1x1_int8.c:56),
 * shl_rvv_gemm_4x16_int8 under shl_rvv_conv_im2col_gemm_int8
 * (source/thead_rvv/int8/convolution_gemm_int8.c:173, gemm_int8.c:37),
 * shl_rvv_fullyconnected_int8 (source/thead_rvv/int8/fullyconnected_int8.c:94) and their
 * fp16 twins; semantics follow shl_ref_conv2d_quant (source/reference/convolution.c:370)
 * and shl_ref_fullyconnected_quant (source/reference/fullyconnected.c:54). */
typedef struct {
    int32_t dtype;   /* B200_I8 | B200_F16 */
    int32_t m, n, k; /* logical sizes: rows, output channels, reduction length      */
    const void *a;   /* device [m][lda]                                             */
    int32_t lda;     /* elements, lda*elem % 16 == 0                                */
    const void *w;   /* device [n][ldw]   (OIHW with H=W=1, or fc weight [O][I])    */
    int32_t ldw;     /* elements, ldw*elem % 16 == 0                                */
    void *out;       /* device [m][ldo]; columns [n, ldo) hold unspecified values    */
    int32_t ldo;     /* elements, ldo*elem % 16 == 0, ldo >= n                      */
    int32_t out_cols; /* 0, or the number of columns of a row this call may write, starting at `out`
                         (n <= out_cols <= ldo, out_cols*elem % 16 == 0): a group of a grouped convolution
                         writes its own column window of the shared output, so whole tiles must be clipped
                         there and not at the row pitch                                              */
    b200_epilogue ep;
    /* asymmetric weights (both NULL for symmetric ones): per-output-channel weight zero points and the
     * row sums rs[m] = sum_k a[m][k] - zp_in * k computed by b200_rowsum_i8 */
    const int32_t *w_zp;   /* device [n] */
    const int32_t *rowsum; /* device [m] */
    int32_t w_dynamic;     /* != 0: `w` was written by the preceding kernel of the stream (matmul of two activations), not at
                              init: under programmatic dependent launch the kernel must not prefetch it before its
                              predecessor has completed */
} b200_gemm_desc;
int b200_gemm(const b200_gemm_desc *d, void *stream);
/* rs[m] = sum_{k < K} a[m][k] - zp_in * K over int8 rows of pitch lda: the per-pixel term an asymmetric
 * weight zero point multiplies (see the contract above).  On an im2col matrix the padded taps hold zp_in
 * and cancel, as they must. */
int b200_rowsum_i8(const void *a, int32_t lda, int32_t m, int32_t k, int32_t zp_in, int32_t *rs, void *stream);

/* ---- implicit-GEMM conv2d: k > 1 convolutions without an im2col matrix ---------------------- */
/* out[(b, oy, ox)][o] = epilogue( sum over taps and channels of x~ * w ): the same contract and packed weights
 * ([o][ldw], k = (ky, kx, ci)) as b200_im2col + b200_gemm, but the GEMM's A operand is gathered by TMA in im2col
 * mode (cuTensorMapEncodeIm2col: one load = 128 output pixels x one filter tap x one channel slab) straight
 * from the pixel-major activation tensor -- the [M][K] matrix never exists in HBM.  TMA zero-fills padded taps
 * where the contract wants zp_in; the difference enters through per-border-class accumulator seeds:
 *   cls_map[oy * ow + ox] = class of the output position (which taps fall into the padding),
 *   seeds[cls][o]         = ibias[o] + zp_in * sum of w[o] over those taps (all channels)
 * (ncls <= 1: no correction -- zp_in == 0 or no padding; ep.ibias is used as it is).
 * int8, group 1, channels a multiple of 64.  Replaces shl_rvv_conv_im2col_gemm_int8
 * (source/thead_rvv/int8/convolution_gemm_int8.c:106-170) without its im2col buffer. */
typedef struct {
    int32_t n, h, w, c, cp_in;  /* input [n][h][w][cp_in], c channels                         */
    int32_t o, oh, ow;
    int32_t kh, kw, stride_h, stride_w, pad_top, pad_left, dil_h, dil_w;
    const void *in;
    const void *wt;             /* device [o][ldw] int8                                       */
    int32_t ldw;
    void *out;                  /* device [n][oh][ow][ldo]                                    */
    int32_t ldo;
    b200_epilogue ep;
    int32_t ncls;
    const int32_t *seeds;       /* device [ncls][o]                                           */
    const uint8_t *cls_map;     /* device [oh * ow]                                           */
    int32_t dw_slab;            /* != 0: DEPTHWISE convolution (o == c) on the same kernel: every 64-channel n-tile
                                   contracts over taps x its own 64 input channels against a B that is diagonal per tap --
                                   wt = [o][kh*kw*64] int8, zero except w[o][tap] at column tap*64 + o % 64.  The tensor
                                   pipe spends 64x the useful MACs, and still beats the dp4a kernels: what a depthwise
                                   output costs on the CUDA cores is its requantise epilogue plus 13 instructions of tap
                                   transposition and dp4a, here only the epilogue is left */
} b200_conv_igemm_desc;
int b200_conv_igemm_supported(const b200_conv_igemm_desc *d);
int b200_conv_igemm(const b200_conv_igemm_desc *d, void *stream);

/* ---- im2col (the other half of "im2col + GEMM conv2d") ------------------------- */
/* col[m][k], m = (b, oy, ox), k = (ky, kx, ci) with ci fastest, row pitch ldk elements;
 * padded taps and k >= kh*kw*cg are written as `pad_value` (= zp_in for int8) / 0.
 * in_nchw != 0: `in` is the API-side NCHW tensor (a network's first layer reads it
 * directly); otherwise pixel-major with channel stride cp_in.  `c_off` selects the
 * first channel of a group.  Replaces the im2col loop of
 * shl_rvv_conv_im2col_gemm_int8 (convolution_gemm_int8.c:106-134) and
 * conv_im2col_sgemm_avx's (source/reference/conv_avx.h:109). */
typedef struct {
    int32_t dtype;
    int32_t n, h, w, cp_in, in_nchw, c_total; /* c_total: channels of `in` (NCHW indexing) */
    int32_t c_off, cg;                         /* channel window gathered                  */
    int32_t oh, ow, kh, kw, stride_h, stride_w, pad_top, pad_left, dil_h, dil_w;
    int32_t ldk;
    int32_t pad_value;
    const void *in;
    void *col;
} b200_im2col_desc;
int b200_im2col(const b200_im2col_desc *d, void *stream);

/* ---- direct conv2d for a first layer (int8, few input channels, NCHW input) ---- */
/* The one conv shape where im2col + GEMM loses: K = c*kh*kw <= 160 (3x3x3, 7x7x3 stems) read
 * straight from the API's NCHW tensor, dp4a against weights in shared memory, pixel-major
 * output.  Same semantics and epilogue as b200_gemm on the im2col matrix; `wt` is the same packed
 * [o][ldw] (k = (ky, kx, c)) buffer the GEMM path uses.  Replaces, for this shape,
 * shl_rvv_conv_im2col_gemm_int8 (source/thead_rvv/int8/convolution_gemm_int8.c:173). */
typedef struct {
    int32_t n, c, h, w;      /* NCHW input                                        */
    int32_t o, oh, ow, cp_out;
    int32_t kh, kw, stride_h, stride_w, pad_top, pad_left, dil_h, dil_w;
    int32_t ldw;             /* weight row pitch, bytes                            */
    const void *in;
    const void *wt;
    void *out;               /* device [n][oh][ow][cp_out]                         */
    int32_t zp_in;
    b200_epilogue ep;
} b200_conv_direct_desc;
int b200_conv2d_direct(const b200_conv_direct_desc *d, void *stream);
/* fp16 twin (same descriptor; f32 accumulation; C*kh*kw <= 160, O <= 64) */
int b200_conv2d_direct_f16(const b200_conv_direct_desc *d, void *stream);

/* ---- depthwise conv2d (HBM-bound stencil) ------------------------------------ */
/* Replaces shl_rvv_dwconv3x3s1_int8 / s2 (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31 )
 * and the fp16 twins; semantics: shl_ref_depthwise_conv2d_quant (source/reference/convolution.c:416).
 * Weights are packed once at init from O1HW (layouts below). depth_multiplier == 1. */
typedef struct {
    int32_t dtype;
    int32_t n, c, cp; /* batch, channels, channel stride (elements)          */
    int32_t h, w, oh, ow;
    int32_t kh, kw, stride_h, stride_w, pad_top, pad_left, dil_h, dil_w;
    const void *in;  /* device [n][h][w][cp]                                */
    const void *wt;  /* device, tap-major [kh*kw][cp]: f16 halves, or for int8 one 32-bit word
                        per (tap, channel) with the weight in byte lane (c & 3), others 0  */
    const void *wt_row3; /* optional, int8 3x3 only: [3 (ky)][cp] words
                            (w[ky][0][c], w[ky][1][c], w[ky][2][c], 0) for the TMA-fed dp4a kernel */
    void *out;       /* device [n][oh][ow][cp]                              */
    int32_t zp_in;   /* int8: value of a padded tap                         */
    b200_epilogue ep;
    const int32_t *w_zp; /* device [cp] weight zero points, or NULL (symmetric weights); with them ep.ibias[o] must be
                            -zp_in * (sum_taps w[o] - taps * w_zp[o]) and the generic kernel is used */
} b200_dwconv_desc;
int b200_dwconv2d(const b200_dwconv_desc *d, void *stream);

/* ---- depthwise 3x3 -> pointwise 1x1 in one kernel (a MobileNet block) -------------- */
/* out = pw(dw(in)): the int8 result of the depthwise stage (requantised with dw.ep exactly as
 * b200_dwconv2d would) is written into the tcgen05 GEMM's swizzled A operand tile in shared memory
 * and never reaches HBM; the pointwise stage is b200_gemm's contract with `ep`.  Bit-identical to
 * b200_dwconv2d followed by b200_gemm.  Covers int8, 3x3, stride 1 / 2, pads <= 1, o <= 256 and
 * pointwise weights that stay resident in shared memory (b200_dwpw_supported tells).  dw.out is
 * ignored.  Replaces the pair shl_rvv_dwconv3x3s1_int8 / s2
 * (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31,244) -> shl_rvv_conv1x1s1_gemm_int8
 * (source/thead_rvv/int8/convolution_1x1_int8.c:56) of example/c906_mobilenetv1_f16.c. */
typedef struct {
    b200_dwconv_desc dw;
    int32_t o;       /* pointwise output channels                                     */
    const void *w;   /* device [o][ldw] int8, the b200_gemm weight layout (K = dw.c)  */
    int32_t ldw;
    void *out;       /* device [n][oh][ow][ldo]                                       */
    int32_t ldo;
    b200_epilogue ep; /* pointwise epilogue                                           */
} b200_dwpw_desc;
int b200_dwpw_supported(const b200_dwpw_desc *d);
int b200_dwpw_fused(const b200_dwpw_desc *d, void *stream);
/* the tile plan the fused kernel would use, as text (0 = the pair is not covered); tools and tests */
int b200_dwpw_plan_describe(const b200_dwpw_desc *d, char *buf, int buflen);

/* ---- bandwidth-bound ops ------------------------------------------------------ */
/* q' = lut[q + 128] over `count` bytes: relu / relu6 / requantising identity with per-tensor
 * qinfo (source/reference/relu.c:39, relu6.c:42); count % 16 == 0 on pixel-major tensors. */
int b200_lut_i8(const void *in, void *out, size_t count, const int8_t *lut_dev, void *stream);
/* fp16 relu / relu6 */
/* fp16 unary ops of b200_act (f32 arithmetic on the converted value, like the reference's fp16 path) */
int b200_unary_f16(const void *in, void *out, size_t count, int act, float p0, float p1, void *stream);
int b200_relu_f16(const void *in, void *out, size_t count, int act, void *stream);
/* elementwise add, same shapes, per-tensor qinfo (source/reference/add.c:36 through
 * diso_callback_base utils.c:622): r = (qa-zpa)*sa + (qb-zpb)*sb ; q = quant(r) ; optional
 * fused relu expressed as a post table (as in b200_epilogue). */
/* elementwise binary ops between tensors of the same shape (shl_rvv_add/sub/mul_int8; semantics
 * source/reference/add.c:36, sub.c:36, mul.c:36 through shl_ref_diso_callback_base) */
typedef enum {
    B200_BINOP_ADD = 0,
    B200_BINOP_SUB = 1,
    B200_BINOP_MUL = 2,
    B200_BINOP_PRELU = 3, /* a >= 0 ? a : a * b, b = the per-channel slope (source/reference/prelu.c:44-48) */
    B200_BINOP_DIV = 4    /* a / b (source/reference/div.c:22); int8: +-inf saturate, 0 / 0 gives 0 -- what the
                             reference's (int8_t)NaN yields on its x86 build (source/nn2/utils.c:550-560) */
} b200_binop;
int b200_binary(int binop, int dtype, const void *a, const void *b, void *out, size_t count, float s_a, int zp_a,
                float s_b, int zp_b, float s_out, int zp_out, const int8_t *post_lut, int act, void *stream);
/* the same with a second operand that repeats every `b_count` elements (0 = same shape): a per-channel or
 * scalar constant laid out as one pixel's channels ([cp] elements, padding lanes 0) -- the broadcasting
 * shl_ref_add_f32 / sub / mul do for a [1, C, 1, 1] or one-element operand (source/reference/add.c:21) */
int b200_binary_bcast(int binop, int dtype, const void *a, const void *b, size_t b_count, void *out, size_t count,
                      float s_a, int zp_a, float s_b, int zp_b, float s_out, int zp_out, const int8_t *post_lut, int act,
                      void *stream);
/* the same with one instance of the second operand's period per `a_per_instance` elements of the first: an
 * activation of shape [N, C, 1, 1] against [N, C, H, W] (a_per_instance = H * W * cp, b_count = cp) -- the per-image,
 * per-channel case of shl_ref_diso_broadcast_base (source/reference/utils.c:83), e.g. a squeeze-and-excitation scale */
int b200_binary_bcast_nc(int binop, int dtype, const void *a, const void *b, size_t b_count, size_t a_per_instance, void *out,
                         size_t count, float s_a, int zp_a, float s_b, int zp_b, float s_out, int zp_out,
                         const int8_t *post_lut, int act, void *stream);
int b200_add(int dtype, const void *a, const void *b, void *out, size_t count, float s_a, int zp_a,
             float s_b, int zp_b, float s_out, int zp_out, const int8_t *post_lut, int act,
             void *stream);

/* concat (source/reference/concat.c:20-72; replaces shl_rvv_concat_int8 / _fp16 of
 * source/thead_rvv/setup.c): ONE input copied into its slice of the pixel-major output.  axis counts
 * in (n, c, h, w) = 0..3, `offset` = sum of the earlier inputs' extents along it.  int8: `lut`
 * (device, 256 bytes, index q + 128) is requant_out(dequant_in(q)) for this input; NULL = plain
 * copy (fp16, where f16 -> f32 -> f16 is the identity).  The output's padding lanes are zeroed by
 * the slice that ends at the last channel. */
typedef struct {
    int32_t dtype;
    int32_t n, c, h, w, cp_in;      /* this input */
    int32_t on, oc, oh, ow, cp_out; /* the whole output */
    int32_t axis, offset;
    const void *in;
    void *out;
    const int8_t *lut;
    int32_t extract; /* 0: `in` is the slice (n, c, h, w, cp_in), `out` the whole tensor (on .. cp_out) -- concat;
                        1: `in` is the whole tensor (on, oc, oh, ow, cp_in), `out` the slice (n, c, h, w, cp_out)
                        -- split (source/reference/split.c:20), lut = requant_slice(dequant_whole(q)) */
} b200_concat_desc;
int b200_concat_slice(const b200_concat_desc *d, void *stream);

/* ---- structural / normalisation operators of the RVV table (csrc/tensor_ops.cu) ---------------------------- */
/* A tensor of logical rank 1..4 in the device layout: rank 4 (d0..d3) = (n, c, h, w), rank 3 = (n, c, w), rank 2 =
 * (n, c), rank 1 = (c); element (n, c, h, w) at ((n*H + h)*W + w)*cp + c.  The kernels below walk LOGICAL row-major
 * indices -- the order the reference's loops define the semantics in -- and map them to device offsets. */
typedef struct {
    int32_t rank;
    int32_t dim[4];
    int32_t cp; /* channel pitch in elements */
} b200_view;
/* transpose: out[o] = in[i] with i[perm[k]] = o[k] (source/reference/transpose.c:57; replaces shl_rvv_transpose_int8 /
 * _fp16).  int8 with differing qinfo: lut = requant_out(dequant_in(q)) as for concat, else NULL. */
int b200_permute(const b200_view *in, const void *src, const b200_view *out, void *dst, const int32_t *perm,
                 int elem_bytes, const int8_t *lut_dev, void *stream);
/* gather along `axis` with n_idx constant indices (negative = from the end, out of range = 0.0:
 * source/reference/gather.c:21-60; replaces shl_rvv_gather_int8 / _fp16); oob_q = the quantised 0.0 */
int b200_gather(const b200_view *in, const void *src, const b200_view *out, void *dst, int axis, const int32_t *idx_dev,
                int n_idx, int elem_bytes, const int8_t *lut_dev, int oob_q, void *stream);
/* sum over one axis (axis < 0: over everything), sequential f32 in index order like source/reference/reduce_sum.c:21
 * (replaces shl_rvv_reduce_sum_int8) */
int b200_reduce_sum(const b200_view *in, const void *src, const b200_view *out, void *dst, int axis, int elem_bytes,
                    float s_in, int zp_in, float s_out, int zp_out, void *stream);
/* layer_norm (rms = 0: (x - mean) / sqrt(var + eps) * gamma + beta) / rms_norm (rms = 1: x / sqrt(mean(x^2) + eps) *
 * gamma) over the axes [axis, rank): source/reference/layer_norm.c:21, rms_norm.c:21, the float sequence verbatim;
 * gamma / beta = dequantised f32 device arrays of the normalised extent (replaces shl_rvv_layer_norm_int8 / _fp16,
 * shl_rvv_rms_norm_fp16) */
int b200_norm(int rms, const b200_view *v, const void *src, void *dst, int axis, float eps, const float *gamma_dev,
              const float *beta_dev, int elem_bytes, float s_in, int zp_in, float s_out, int zp_out, void *stream);
/* csinn_matmul on the tcgen05 GEMM: logical matrices [.., R, K] (trans = 0) or [.., K, R] (trans = 1) <-> dense K-major
 * rows [(batch * R + r)][ld] that b200_gemm takes as A or W; and the GEMM's output rows back into a tensor */
int b200_pack_rows(const b200_view *t, const void *src, int batches, int rows, int k, int trans, void *rows_dev, int ld,
                   int elem_bytes, void *stream);
int b200_unpack_rows(const b200_view *t, void *dst, long long nrows, int cols, const void *rows_dev, int ld, int elem_bytes,
                     void *stream);

/* maxpool / avgpool on pixel-major tensors; reference: source/reference/maxpool.c:64,
 * averagepool.c:71 (sequential f32 sum in (y,x) order, divided by the valid-tap count unless
 * count_include_pad). global average pool = kh=h, kw=w (global_averagepool.c:21). */
typedef struct {
    int32_t dtype;
    int32_t n, c, cp, h, w, oh, ow;
    int32_t kh, kw, stride_h, stride_w, pad_top, pad_left;
    int32_t is_avg, count_include_pad;
    float s_in;
    int32_t zp_in;
    float s_out;
    int32_t zp_out;
    const void *in;
    void *out;
} b200_pool_desc;
int b200_pool2d(const b200_pool_desc *d, void *stream);

/* softmax over the channel axis of [rows][cp] (axis=1 of an NCHW tensor with H=W=1);
 * source/reference/softmax.c:20-66 (double exp, f32 accumulate in channel order). */
/* test hook (csrc/umma_probe.cu): one M128 N32 K32 kind::i8 MMA whose SWIZZLE_128B A descriptor starts `shift`
 * 128-byte rows into a TMA-written tile: out[m][n] = sum_k a[shift + m][32 k0 + k] * b[n][32 k0 + k].
 * a_dev [144][128] int8, b_dev [32][128] int8, out_dev [128][32] int32 */
int b200_test_umma_shifted_start(const void *a_dev, const void *b_dev, int shift, int k0, void *out_dev, void *stream);
/* test hook (csrc/umma_probe.cu): cycles one SM needs for `reps` back-to-back M128 x n x K32 kind::i8 MMAs with
 * `nacc` accumulators in rotation (1 = one dependent chain) */
int b200_test_umma_rate(int n, int nacc, int reps, long long *cycles_host, void *stream);
int b200_test_umma_rate2(int m, int n, int f16, int nacc, int reps, long long *cycles_host, void *stream);
int b200_test_umma_rate3(int m, int n, int f16, int nacc, int reps, int issuers, long long *cycles_host, void *stream);
/* test hook (csrc/umma_probe.cu): one TMA im2col-mode load (cuTensorMapEncodeIm2col map over [n][h][w][cp] uint8) of
 * `pixels` output pixels x `chans` channels of the tap at (woff, hoff), starting at base pixel (w0, h0, n0), channel c0;
 * out_dev receives the [pixels][chans] bytes as they lie in shared memory (no swizzle) */
int b200_test_tma_im2col(const void *in_dev, int n, int h, int w, int c, int cp, int lower_w, int lower_h, int upper_w,
                         int upper_h, int stride_w, int stride_h, int chans, int pixels, int c0, int w0, int h0, int n0, int woff,
                         int hoff, void *out_dev, void *stream);
/* test hook: the softmax denominator code alone (rows x c doubles -> rows floats), see csrc/softmax.cu */
int b200_test_softmax_denominator(const void *e_dev, int rows, int c, void *out_dev, void *stream);
int b200_softmax(int dtype, const void *in, void *out, int rows, int c, int cp_in, int cp_out,
                 float s_in, int zp_in, float s_out, int zp_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NN_H_ */
