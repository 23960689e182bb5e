/*
 * oracle_int.h -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * executed from the product (csi-nn2_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liboracle_int.so.
 *
 * CPU restatement (plain C, NCHW, scalar loops) of the quantised hot path of
 * the reference: csinn_conv2d / csinn_depthwise_conv2d / csinn_fullyconnected
 * and the bandwidth ops around them.  Paths cited are relative to the
 * reference tree.
 *
 * Parity pin: tests/test_oracle.py checks every function here against the
 * UNMODIFIED reference compiled from its own sources (oracle/_ref/libshl_ref_x86.so,
 * built by oracle/Makefile) and against the reference's own golden vectors
 * (tests/unit_test/valid_data/{conv2d,dwconv2d,fullyconnected,...}.dat, committed
 * as tests/golden/*.npz by tests/golden/make_golden.py).
 *
 * Arithmetic contract of the contraction ops (conv / depthwise / fc), int8:
 *     acc   = sum over ALL kernel taps of x~ * w          (int32, exact)
 *             x~ = x_q inside the image, zp_in at padded positions
 *             (pad = 0.0 in the real domain, source/reference/convolution.c:62-66,
 *              250-257; the RVV im2col writes zp_in at pads too,
 *              source/thead_rvv/int8/convolution_gemm_int8.c:116,125)
 *     acc  += ibias[o] = -zp_in * sum_taps w[o]            (zero-point fold)
 *     acc  -= zp_w[o] * (sum_taps x~ - zp_in * taps)       (only with asymmetric weights, see zp_w below)
 *     f     = fmaf((float)acc, mult[o], badd[o])           (one rounding)
 *             mult[o] = (float)((double)s_in * s_w[o] / s_out)
 *             badd[o] = (float)((double)bias_q[o] * s_b[o] / s_out)   (s_b defaults to the float
 *                       product s_in * s_w[o], the value a float qinfo field would hold)
 *     q     = clamp((int)rintf(f) + zp_out, -128, 127)     (round half even, as nearbyint in
 *                                                           source/nn2/utils.c:550 float_to_int8_base)
 * The reference itself accumulates dequantised f32 products
 * (source/reference/utils.c:639-655), so its result carries f32 rounding noise:
 * it differs from this exact-integer restatement by +-1 LSB on a few ppm of
 * outputs (measured in tests/test_oracle.py; its own AVX and non-AVX builds
 * disagree with each other at the same rate).
 *
 * Elementwise / pooling ops restate the reference's float sequence exactly
 * (dequantise -> f32 op -> requantise, source/reference/utils.c:609-637) and are
 * bit-exact against the reference library.
 */
#ifndef ORACLE_INT_H_
#define ORACLE_INT_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_ACT_NONE = 0, ORACLE_ACT_RELU = 1, ORACLE_ACT_RELU6 = 2 };

typedef struct {
    int32_t n, c, h, w;          /* input NCHW                                   */
    int32_t o, kh, kw;           /* kernel OIHW, I = c / group                    */
    int32_t oh, ow;
    int32_t stride_h, stride_w, pad_top, pad_left, dil_h, dil_w, group;
    /* quantisation */
    float s_in;
    int32_t zp_in;
    const float *s_w;            /* [o] (per-channel) or [1] when w_channels == 1 */
    int32_t w_channels;
    const float *s_b;            /* bias scales, same indexing as s_w; NULL -> s_in*s_w */
    float s_out;
    int32_t zp_out;
    int32_t fuse_zp2bias;        /* bias already holds -zp_in*sum(w) (thead_rvv/int8/convolution.c:172) */
    /* fused activation (CSINN_OP_CONV2D_RELU / _RELU6: convolution_relu.c:34, convolution_relu6.c:21) */
    int32_t act;
    /* optional second stage = a standalone relu/relu6 node with its own qinfo */
    int32_t post;
    int32_t post_act;
    float post_s_out;
    int32_t post_zp_out;
    /* weight zero points, indexed like s_w, or NULL for symmetric weights (zero_point 0).  With them
     *     acc = sum over taps of (x~ - zp_in) * (w - zp_w[o])
     *         = sum x~ * w + ibias[o] - zp_w[o] * (sum_taps x~ - zp_in * taps)
     * (a padded tap holds zp_in and contributes nothing); the reference dequantises the kernel with its
     * zero point, source/nn2/utils.c:920-931 nchw_int8_to_float -> int8_to_float_base */
    const int32_t *zp_w;
} oracle_conv_params;

/* requant tables shared by conv / dw / fc; returns 0, or -1 if a bound check fails */
int oracle_requant_tables(const oracle_conv_params *p, const int8_t *wt, const int32_t *bias,
                          int taps_per_o, float *mult, float *badd, int32_t *ibias);

/* csinn_conv2d / group conv, int8 (source/reference/convolution.c:370 shl_ref_conv2d_quant) */
int oracle_conv2d_i8(const oracle_conv_params *p, const int8_t *in, const int8_t *wt,
                     const int32_t *bias, int8_t *out);
/* csinn_depthwise_conv2d int8 (convolution.c:416); kernel layout O1HW, depth multiplier o/c */
int oracle_dwconv2d_i8(const oracle_conv_params *p, const int8_t *in, const int8_t *wt,
                       const int32_t *bias, int8_t *out);
/* csinn_fullyconnected int8 (source/reference/fullyconnected.c:54): in [n][c], wt [o][c] */
int oracle_fc_i8(const oracle_conv_params *p, const int8_t *in, const int8_t *wt,
                 const int32_t *bias, int8_t *out);

/* float restatements used for the fp16 path (f32 accumulate like the reference; inputs are
 * already-dequantised floats; convolution.c:28-89 / 206-269 / fullyconnected.c:21-52) */
int oracle_conv2d_f32(const oracle_conv_params *p, const float *in, const float *wt,
                      const float *bias, float *out);
int oracle_dwconv2d_f32(const oracle_conv_params *p, const float *in, const float *wt,
                        const float *bias, float *out);
int oracle_fc_f32(const oracle_conv_params *p, const float *in, const float *wt,
                  const float *bias, float *out);

/* elementwise, bit-exact float sequence (relu.c:39, relu6.c, add.c:36) */
void oracle_relu_i8(const int8_t *in, int8_t *out, int64_t count, int act, float s_in, int zp_in,
                    float s_out, int zp_out);
enum { ORACLE_UNARY_LEAKY_RELU = 3, ORACLE_UNARY_SIGMOID = 4, ORACLE_UNARY_CLIP = 5, ORACLE_UNARY_SILU = 6, ORACLE_UNARY_ERF = 7 };
void oracle_unary_i8(const int8_t *in, int8_t *out, int64_t count, int op, float p0, float p1, float s_in,
                     int zp_in, float s_out, int zp_out);
void oracle_add_i8(const int8_t *a, const int8_t *b, int8_t *out, int64_t count, float s_a,
                   int zp_a, float s_b, int zp_b, float s_out, int zp_out);

/* op: 0 add, 1 sub, 2 mul */
void oracle_concat_i8(int k, const int8_t *const *in, const int64_t *axis_dim, const float *s_in,
                      const int32_t *zp_in, int64_t outer, int64_t inner, float s_out, int zp_out, int8_t *out);
void oracle_binary_i8(int op, const int8_t *a, const int8_t *b, int8_t *out, int64_t count, float s_a,
                      int zp_a, float s_b, int zp_b, float s_out, int zp_out);

typedef struct {
    int32_t n, c, h, w, oh, ow, kh, kw, stride_h, stride_w, pad_top, pad_left;
    int32_t count_include_pad;
    float s_in;
    int32_t zp_in;
    float s_out;
    int32_t zp_out;
} oracle_pool_params;
/* averagepool.c:71, maxpool.c:64, global_averagepool.c:21 (kh=h, kw=w) */
void oracle_avgpool_i8(const oracle_pool_params *p, const int8_t *in, int8_t *out);
void oracle_maxpool_i8(const oracle_pool_params *p, const int8_t *in, int8_t *out);
/* softmax.c:20 over axis 1 of [rows][c] */
void oracle_softmax_i8(const int8_t *in, int8_t *out, int rows, int c, float s_in, int zp_in,
                       float s_out, int zp_out);

/* fp16 <-> f32 exactly as the reference converts (source/nn2/utils.c:576-660) */
uint16_t oracle_f32_to_f16(float v);
float oracle_f16_to_f32(uint16_t v);

#ifdef __cplusplus
}
#endif
#endif
