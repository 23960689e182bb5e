/*
 * quant.c -- init-time host work of the b200 backend: per-channel requantisation tables and
 * weight packing.  Does what shl_rvv_conv2d_init_int8 does at init
 * (source/thead_rvv/int8/convolution.c:161-190: per-channel multiplier, zero-point fold) and
 * what shl_rvv_conv_im2col_gemm_reorder_kernel_int8 / shl_rvv_fc_gemm_reorder_weight_int8 do
 * (convolution_gemm_int8.c:21, fullyconnected_int8.c:79), but for the epilogue contract of
 * include/b200nn.h and the K-major tile layout the tcgen05 GEMM reads; unlike the RVV back end
 * it never mutates the caller's kernel / bias buffers.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "b200_internal.h"

float b200_f16_to_f32(uint16_t h);
static float f16_to_f32(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF, u;
    if (exp == 0) {
        if (man == 0) {
            u = sign;
        } else {
            int e = -1;
            do {
                man <<= 1;
                e++;
            } while (!(man & 0x400));
            u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
        }
    } else if (exp == 31) {
        u = sign | 0x7F800000u | (man << 13);
    } else {
        u = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &u, 4);
    return f;
}

float b200_f16_to_f32(uint16_t h) { return f16_to_f32(h); }

/* IEEE round-to-nearest-even, what a hardware convert does */
static uint16_t f32_to_f16_rne(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const int32_t exp = (int32_t)((u >> 23) & 0xFF) - 127 + 15;
    uint32_t man = u & 0x7FFFFFu;
    if (((u >> 23) & 0xFF) == 0xFF) return (uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7C00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        const int shift = 14 - exp;
        uint32_t half = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13);
    const uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++; /* a carry into the exponent is the right answer */
    return (uint16_t)(sign | half);
}

/* CSINN_QUANT_FLOAT16_W_INT8 (weight-only quantisation under fp16 activations): the int8 kernel is
 * dequantised to fp16 once at init -- (q - zp) * scale with the per-channel qinfo of output channel
 * dim[0], the float sequence of the reference's kernel transform (source/nn2/utils.c:920-931) -- as
 * shl_rvv_conv_im2col_gemm_dequantize_per_channel_i8_to_f16 does for the C906 (c906_opt/fp16/convolution.c:77-81).
 * Returns a heap copy of the tensor header with fp16 data; release with b200_free_dequant. */
struct csinn_tensor *b200_dequant_weights_f16(const struct csinn_tensor *kernel)
{
    if (!kernel->data || !kernel->qinfo || kernel->dim_count < 1) {
        b200_fail("int8 weights under fp16 activations need data and qinfo");
        return NULL;
    }
    int64_t total = 1;
    for (int i = 0; i < kernel->dim_count; i++) total *= kernel->dim[i];
    const int64_t per_o = total / kernel->dim[0];
    struct csinn_tensor *t = malloc(sizeof(*t));
    uint16_t *h = malloc((size_t)total * sizeof(uint16_t));
    if (!t || !h) {
        free(t);
        free(h);
        b200_fail("out of host memory dequantising weights");
        return NULL;
    }
    memcpy(t, kernel, sizeof(*t));
    const int8_t *q = kernel->data;
    for (int64_t i = 0; i < total; i++) {
        const int qi = kernel->quant_channel > 1 ? (int)(i / per_o) : 0;
        const float v = ((float)q[i] - (float)kernel->qinfo[qi].zero_point) * kernel->qinfo[qi].scale;
        h[i] = f32_to_f16_rne(v);
    }
    t->dtype = CSINN_DTYPE_FLOAT16;
    t->data = h;
    return t;
}
/* The same tensor as EXACT integers in fp16 (q - zp, |v| <= 255) with the per-channel scales handed to the next
 * b200_make_requant call, which uploads them as the GEMM epilogue's per-column multiplier: products of fp16
 * activations with integer weights are exact in the f32 accumulator and the scale is applied once, in f32 --
 * what the reference's f32 path computes up to summation order -- instead of rounding every weight to fp16 first
 * (5e-4 relative each).  Used for the tcgen05 GEMM path; depthwise and first-layer kernels keep dequantised weights. */
static float *g_pending_wscale;
static int g_pending_wscale_n;
struct csinn_tensor *b200_int8_weights_as_f16(const struct csinn_tensor *kernel)
{
    struct csinn_tensor *t = b200_dequant_weights_f16(kernel);
    if (!t) return NULL;
    int64_t total = 1;
    for (int i = 0; i < kernel->dim_count; i++) total *= kernel->dim[i];
    const int O = kernel->dim[0];
    const int64_t per_o = total / O;
    float *sc = malloc((size_t)O * sizeof(float));
    if (!sc) {
        b200_free_dequant(t);
        return NULL;
    }
    const int8_t *q = kernel->data;
    uint16_t *h = t->data;
    for (int64_t i = 0; i < total; i++) {
        const int qi = kernel->quant_channel > 1 ? (int)(i / per_o) : 0;
        h[i] = f32_to_f16_rne((float)q[i] - (float)kernel->qinfo[qi].zero_point);
    }
    for (int o = 0; o < O; o++) sc[o] = kernel->qinfo[kernel->quant_channel > 1 ? o : 0].scale;
    free(g_pending_wscale);
    g_pending_wscale = sc, g_pending_wscale_n = O;
    return t;
}
void b200_free_dequant(struct csinn_tensor *t)
{
    if (!t) return;
    free(t->data);
    free(t);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static int has_bias(const struct csinn_tensor *bias)
{
    return bias && bias->data && bias->dim_count != 0 && csinn_tensor_size((struct csinn_tensor *)bias) != 0;
}

int b200_make_requant(b200_op *op, const struct csinn_tensor *input,
                      const struct csinn_tensor *kernel, const struct csinn_tensor *bias,
                      const struct csinn_tensor *output, int taps_per_o, int fuse_zp2bias,
                      int n_out)
{
    const int n_alloc = round_up(n_out, 16);
    float *mult = calloc(n_alloc, sizeof(float));
    float *badd = calloc(n_alloc, sizeof(float));
    int32_t *ibias = calloc(n_alloc, sizeof(int32_t));
    int32_t *wzp = calloc(n_alloc, sizeof(int32_t));
    int any_wzp = 0;
    int rc = CSINN_TRUE;
    if (!mult || !badd || !ibias || !wzp) {
        b200_fail("out of host memory building requant tables");
        rc = CSINN_FALSE;
        goto done;
    }
    if (op->dtype == B200_F16) {
        if (has_bias(bias)) {
            const uint16_t *b = bias->data;
            /* the reference scales fp16 constants by qinfo->scale when it is not 1 (f16_to_float,
             * source/nn2/utils.c:1183-1188) */
            const float bs = bias->qinfo && fabsf(bias->qinfo->scale - 1.f) > 1.1920929e-7f ? bias->qinfo->scale : 1.f;
            for (int o = 0; o < n_out; o++) badd[o] = f16_to_f32(b[o]) * bs;
        }
        op->d_mult = NULL;
        op->d_ibias = NULL;
        op->d_badd = b200_warena_put(op->ctx, badd, n_alloc * sizeof(float));
        if (!op->d_badd) rc = CSINN_FALSE;
        if (g_pending_wscale) { /* integer weights in fp16 (b200_int8_weights_as_f16): per-column scale in the epilogue */
            if (g_pending_wscale_n == n_out) {
                for (int o = 0; o < n_out; o++) mult[o] = g_pending_wscale[o];
                op->d_mult = b200_warena_put(op->ctx, mult, n_alloc * sizeof(float));
                if (!op->d_mult) rc = CSINN_FALSE;
            } else {
                b200_fail("weight scales for %d channels handed to an operator with %d outputs", g_pending_wscale_n, n_out);
                rc = CSINN_FALSE;
            }
            free(g_pending_wscale);
            g_pending_wscale = NULL;
        }
        goto done;
    }
    if (!input->qinfo || !kernel->qinfo || !output->qinfo) {
        b200_fail("int8 op without qinfo on input / kernel / output");
        rc = CSINN_FALSE;
        goto done;
    }
    const double s_in = input->qinfo->scale, s_out = output->qinfo->scale;
    const int zp_in = input->qinfo->zero_point;
    const int8_t *w = kernel->data;
    const int32_t *b = has_bias(bias) ? bias->data : NULL;
    for (int o = 0; o < n_out; o++) {
        const int qi = kernel->quant_channel > 1 ? o : 0;
        /* asymmetric weights: acc = sum (x~ - zp_in) * (w - zw).  Contraction ops subtract zw * (row sum of x~ - zp_in)
         * in the epilogue (include/b200nn.h); the depthwise kernel accumulates the window sum per channel. */
        const int zw = kernel->qinfo[qi].zero_point;
        if (zw < -128 || zw > 127) {
            b200_fail("weight zero_point %d (channel %d) outside int8", zw, o);
            rc = CSINN_FALSE;
            goto done;
        }
        wzp[o] = zw;
        any_wzp |= zw != 0;
        const double sw = kernel->qinfo[qi].scale;
        /* default bias scale: the float product, as it would sit in bias->qinfo[].scale */
        const float sb_default = input->qinfo->scale * kernel->qinfo[qi].scale;
        double sb = sb_default;
        if (b && bias->qinfo) {
            const int bi = bias->quant_channel > 1 ? o : 0;
            if (bias->qinfo[bi].scale != 0) sb = bias->qinfo[bi].scale;
        }
        int64_t wsum = 0, wabs = 0;
        for (int t = 0; t < taps_per_o; t++) {
            const int wv = w[(int64_t)o * taps_per_o + t];
            wsum += wv;
            wabs += wv - zw < 0 ? zw - wv : wv - zw;
        }
        int64_t bq = b ? b[o] : 0;
        /* un-fold, cf. reference/convolution.c:375-395 (it sums the DEQUANTISED kernel, i.e. w - zw) */
        if (fuse_zp2bias) bq += (int64_t)zp_in * (wsum - (int64_t)taps_per_o * zw);
        mult[o] = (float)(s_in * sw / s_out);
        badd[o] = (float)((double)bq * sb / s_out);
        ibias[o] = (int32_t)(-(int64_t)zp_in * wsum);
        /* The epilogues round through the 1.5 * 2^23 magic constant.  That is exact below 2^22, still lands on
         * the right side of the int8 clamp for every larger POSITIVE value (the float's bit pattern only grows)
         * and for negative values down to -1.5 * 2^23, where the sum changes sign and the integer reading of
         * the bits would wrap.  So the requirement is |f| < 2^23, checked with this channel's own weights:
         * |acc - zp_in * wsum| <= max|x - zp_in| * sum|w|, not the data-independent K * 128 * 255 (which refused
         * e.g. a K = 25088 fullyconnected with ordinary scales). */
        const int xmax = (127 - zp_in) > (zp_in + 128) ? (127 - zp_in) : (zp_in + 128);
        const double bound = (double)wabs * xmax * fabs(mult[o]) + fabs(badd[o]);
        if (!(bound < 8388608.0)) {
            b200_fail("channel %d: |acc*mult+bias| may reach %.3g >= 2^23; qinfo out of the supported range",
                      o, bound);
            rc = CSINN_FALSE;
            goto done;
        }
    }
    op->zp_in = zp_in;
    op->zp_out = output->qinfo->zero_point;
    op->s_out = output->qinfo->scale;
    {
        /* quantised 6.0 in the output domain (CONV2D_RELU6: convolution_relu6.c:21) */
        float t = 6.0f / output->qinfo->scale;
        float v = (float)(nearbyint((double)t) + (double)op->zp_out);
        op->q6 = v > 127 ? 127 : (v < -128 ? -128 : (int)v);
    }
    op->d_mult = b200_warena_put(op->ctx, mult, n_alloc * sizeof(float));
    op->d_badd = b200_warena_put(op->ctx, badd, n_alloc * sizeof(float));
    if (any_wzp && op->kind == B200_OPK_DW) /* the depthwise kernel multiplies (x~ - 0) by (w - zw) tap by tap */
        for (int o = 0; o < n_out; o++) ibias[o] += zp_in * taps_per_o * wzp[o];
    op->d_ibias = b200_warena_put(op->ctx, ibias, n_alloc * sizeof(int32_t));
    op->d_wzp = any_wzp ? b200_warena_put(op->ctx, wzp, n_alloc * sizeof(int32_t)) : NULL;
    if (!op->d_mult || !op->d_badd || !op->d_ibias || (any_wzp && !op->d_wzp)) rc = CSINN_FALSE;
done:
    free(mult);
    free(badd);
    free(ibias);
    free(wzp);
    return rc;
}

/* Tables of the implicit-GEMM convolution (include/b200nn.h, b200_conv_igemm_desc): TMA zero-fills the taps that
 * fall outside the image, the contract wants zp_in there.  Output row oy has the kernel rows {ky : oy*sh - pt +
 * ky*dh outside [0, h)} in the padding, output column ox the kernel columns likewise; a position's class is the
 * pair of those two sets, and its accumulator seed -zp_in * (sum of the weights of the taps INSIDE the image). */
int b200_make_igemm_tables(b200_op *op, const struct csinn_tensor *kernel, int h, int w, int oh, int ow)
{
    op->ig_ncls = 0;
    if (op->dtype != B200_I8 || op->group != 1 || op->kh > 16 || op->kw > 16 || (size_t)oh * ow > (1u << 20)) return CSINN_TRUE;
    /* a depthwise kernel is O1HW: one input channel per output */
    const int O = op->o, C = op->kind == B200_OPK_DW ? 1 : op->cin, kh = op->kh, kw = op->kw;
    uint32_t rmasks[64], cmasks[64];
    int nr = 0, nc = 0;
    uint8_t *rcls = malloc((size_t)oh), *ccls = malloc((size_t)ow);
    uint8_t *map = malloc((size_t)oh * ow);
    int rc = CSINN_TRUE;
    if (!rcls || !ccls || !map) goto fail;
    for (int oy = 0; oy < oh; oy++) {
        uint32_t m = 0;
        for (int ky = 0; ky < kh; ky++) {
            const int iy = oy * op->sh - op->pt + ky * op->dh;
            if (iy < 0 || iy >= h) m |= 1u << ky;
        }
        int id = -1;
        for (int i = 0; i < nr; i++)
            if (rmasks[i] == m) id = i;
        if (id < 0) {
            if (nr == 64) goto toomany;
            rmasks[id = nr++] = m;
        }
        rcls[oy] = (uint8_t)id;
    }
    for (int ox = 0; ox < ow; ox++) {
        uint32_t m = 0;
        for (int kx = 0; kx < kw; kx++) {
            const int ix = ox * op->sw - op->pl + kx * op->dw;
            if (ix < 0 || ix >= w) m |= 1u << kx;
        }
        int id = -1;
        for (int i = 0; i < nc; i++)
            if (cmasks[i] == m) id = i;
        if (id < 0) {
            if (nc == 64) goto toomany;
            cmasks[id = nc++] = m;
        }
        ccls[ox] = (uint8_t)id;
    }
    if (nr * nc > 64) goto toomany;
    {
        const int ncls = nr * nc;
        /* interior first would be nicer to read, but any numbering works: class = row class * nc + column class */
        for (int oy = 0; oy < oh; oy++)
            for (int ox = 0; ox < ow; ox++) map[(size_t)oy * ow + ox] = (uint8_t)(rcls[oy] * nc + ccls[ox]);
        int32_t *seeds = calloc((size_t)ncls * O, sizeof(int32_t));
        if (!seeds) goto fail;
        const int8_t *wt = kernel->data; /* OIHW */
        for (int o = 0; o < O; o++) {
            int64_t tap_sum[256];
            for (int t = 0; t < kh * kw; t++) tap_sum[t] = 0;
            for (int ci = 0; ci < C; ci++)
                for (int t = 0; t < kh * kw; t++) tap_sum[t] += wt[((size_t)o * C + ci) * kh * kw + t];
            for (int r = 0; r < nr; r++)
                for (int c = 0; c < nc; c++) {
                    int64_t inside = 0;
                    for (int ky = 0; ky < kh; ky++)
                        for (int kx = 0; kx < kw; kx++)
                            if (!((rmasks[r] >> ky) & 1) && !((cmasks[c] >> kx) & 1)) inside += tap_sum[ky * kw + kx];
                    seeds[(size_t)(r * nc + c) * O + o] = (int32_t)(-(int64_t)op->zp_in * inside);
                }
        }
        op->ig_ncls = (ncls > 1 && op->zp_in != 0) ? ncls : 1;
        op->ig_h = h, op->ig_w = w, op->ig_oh = oh, op->ig_ow = ow;
        if (op->ig_ncls > 1) {
            op->d_ig_seeds = b200_warena_put(op->ctx, seeds, (size_t)ncls * O * sizeof(int32_t));
            op->d_ig_clsmap = b200_warena_put(op->ctx, map, (size_t)oh * ow);
            if (!op->d_ig_seeds || !op->d_ig_clsmap) rc = CSINN_FALSE;
        }
        free(seeds);
    }
    goto done;
toomany:
    op->ig_ncls = 0; /* more than 64 border classes (huge dilated kernels): the explicit im2col path handles it */
    goto done;
fail:
    b200_fail("out of host memory building the implicit-GEMM tables");
    rc = CSINN_FALSE;
done:
    free(rcls);
    free(ccls);
    free(map);
    return rc;
}

/* OIHW -> [O][kh][kw][Cg] rows of pitch ldk: k = (ky, kx, ci), ci fastest = im2col's order */
void *b200_pack_conv_weights(b200_op *op, const struct csinn_tensor *kernel, size_t *bytes)
{
    const int O = kernel->dim[0], cg = kernel->dim[1], kh = kernel->dim[2], kw = kernel->dim[3];
    const int eb = op->eb;
    const size_t total = (size_t)O * op->ldk * eb;
    uint8_t *buf = calloc(1, total ? total : 16);
    if (!buf) return NULL;
    const uint8_t *src = kernel->data;
    for (int o = 0; o < O; o++)
        for (int ci = 0; ci < cg; ci++)
            for (int ky = 0; ky < kh; ky++)
                for (int kx = 0; kx < kw; kx++) {
                    const size_t s = ((((size_t)o * cg + ci) * kh + ky) * kw + kx) * eb;
                    const size_t d = ((size_t)o * op->ldk + ((size_t)ky * kw + kx) * cg + ci) * eb;
                    memcpy(buf + d, src + s, eb);
                }
    void *dev = b200_warena_put(op->ctx, buf, total);
    free(buf);
    *bytes = total;
    return dev;
}

/* O1HW -> tap-major [kh*kw][cp]; int8 entries are expanded to one 32-bit word per channel with
 * the weight in byte lane (c & 3) so that dp4a against a 4-channel activation word yields the
 * single per-channel product */
void *b200_pack_dw_weights(b200_op *op, const struct csinn_tensor *kernel, int cp, size_t *bytes)
{
    const int C = kernel->dim[0], taps = kernel->dim[2] * kernel->dim[3];
    const size_t esz = op->dtype == B200_I8 ? 4 : 2;
    const size_t total = (size_t)taps * cp * esz;
    uint8_t *buf = calloc(1, total ? total : 16);
    if (!buf) return NULL;
    if (op->dtype == B200_I8) {
        const int8_t *src = kernel->data;
        uint32_t *dst = (uint32_t *)buf;
        for (int c = 0; c < C; c++)
            for (int t = 0; t < taps; t++)
                dst[(size_t)t * cp + c] = (uint32_t)(uint8_t)src[(size_t)c * taps + t] << (8 * (c & 3));
    } else {
        const uint16_t *src = kernel->data;
        uint16_t *dst = (uint16_t *)buf;
        for (int c = 0; c < C; c++)
            for (int t = 0; t < taps; t++) dst[(size_t)t * cp + c] = src[(size_t)c * taps + t];
    }
    void *dev = b200_warena_put(op->ctx, buf, total);
    free(buf);
    *bytes = total;
    return dev;
}

/* int8 3x3 depthwise, ky-major for the dp4a kernel: word[ky][c] = (w[ky][0][c], w[ky][1][c],
 * w[ky][2][c], 0) -- one dp4a against the horizontal tap triple of an input row */
void *b200_pack_dw3x3_rows(b200_op *op, const struct csinn_tensor *kernel, int cp)
{
    const int C = kernel->dim[0];
    uint32_t *buf = calloc((size_t)3 * cp, sizeof(uint32_t));
    if (!buf) return NULL;
    const int8_t *src = kernel->data;
    for (int c = 0; c < C; c++)
        for (int ky = 0; ky < 3; ky++) {
            uint32_t w = 0;
            for (int kx = 0; kx < 3; kx++) w |= (uint32_t)(uint8_t)src[(size_t)c * 9 + ky * 3 + kx] << (8 * kx);
            buf[(size_t)ky * cp + c] = w;
        }
    void *dev = b200_warena_put(op->ctx, buf, (size_t)3 * cp * sizeof(uint32_t));
    free(buf);
    return dev;
}

/* depthwise as an implicit GEMM (b200_conv_igemm_desc.dw_slab): row o = [tap][64] bytes, zero except w[o][tap] at
 * column o % 64 of its tap -- a matrix that is diagonal per (64-channel slab, tap) */
void *b200_pack_dw_diag(b200_op *op, const struct csinn_tensor *kernel, int cp, int *ldk)
{
    const int C = kernel->dim[0], taps = kernel->dim[2] * kernel->dim[3];
    const int ld = taps * 64;
    uint8_t *buf = calloc((size_t)cp * ld, 1);
    if (!buf) return NULL;
    const int8_t *src = kernel->data;
    for (int c = 0; c < C; c++)
        for (int t = 0; t < taps; t++) buf[(size_t)c * ld + t * 64 + (c & 63)] = (uint8_t)src[(size_t)c * taps + t];
    void *dev = b200_warena_put(op->ctx, buf, (size_t)cp * ld);
    free(buf);
    *ldk = ld;
    return dev;
}

/* [O][I] -> rows of pitch ldk */
void *b200_pack_fc_weights(b200_op *op, const struct csinn_tensor *weights, size_t *bytes)
{
    const int O = weights->dim[0], I = weights->dim[1];
    const int eb = op->eb;
    const size_t total = (size_t)O * op->ldk * eb;
    uint8_t *buf = calloc(1, total ? total : 16);
    if (!buf) return NULL;
    const uint8_t *src = weights->data;
    for (int o = 0; o < O; o++) memcpy(buf + (size_t)o * op->ldk * eb, src + (size_t)o * I * eb, (size_t)I * eb);
    void *dev = b200_warena_put(op->ctx, buf, total);
    free(buf);
    *bytes = total;
    return dev;
}
