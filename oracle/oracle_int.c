/*
 * oracle_int.c -- TEST INFRASTRUCTURE ONLY (see oracle_int.h for the contract
 * and the parity pin).  Scalar C, NCHW, written for clarity not speed; OpenMP
 * over the outermost loop only so the full-size checks finish in seconds.
 * Compiled with -ffp-contract=off: every float operation below is exactly the
 * one written.
 */
#include "oracle_int.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int clamp_i8(int v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }

/* q = clamp(nearbyint(x / s) + zp): source/nn2/utils.c:550-560 float_to_int8_base.
 * The reference writes `float ret = nearbyint(i / scale) + zp` (float division,
 * double nearbyint, result narrowed to float, then clamped and truncated). */
static inline int8_t quant_i8(float x, float s, int zp)
{
    float r = (float)(nearbyint((double)(x / s)) + (double)zp);
    if (r > 127) return 127;
    if (r < -128) return -128;
    return (int8_t)r;
}
/* (q - zp) * s: source/nn2/utils.c int8_to_float_base */
static inline float dequant_i8(int8_t q, float s, int zp) { return ((float)q - (float)zp) * s; }

/* stage 2 of the epilogue: standalone relu / relu6 node (relu.c:39, relu6.c:42) */
static inline int post_stage(int q, float s_in, int zp_in, int act, float s_out, int zp_out)
{
    float r = dequant_i8((int8_t)q, s_in, zp_in);
    if (act != ORACLE_ACT_NONE) r = r > 0 ? r : 0;
    if (act == ORACLE_ACT_RELU6) r = (float)fmin(r, 6);
    return quant_i8(r, s_out, zp_out);
}

static inline int epilogue_i8(int32_t acc, float mult, float badd, const oracle_conv_params *p)
{
    float f = fmaf((float)acc, mult, badd);
    int q = clamp_i8((int)rintf(f) + p->zp_out);
    if (p->act != ORACLE_ACT_NONE) q = q > p->zp_out ? q : p->zp_out;
    if (p->act == ORACLE_ACT_RELU6) {
        /* conv -> quant -> dequant -> relu6 -> quant with the same qinfo
         * (convolution_relu6.c:21): min(q, quantised 6.0) */
        int q6 = quant_i8(6.0f, p->s_out, p->zp_out);
        q = q < q6 ? q : q6;
    }
    if (p->post) q = post_stage(q, p->s_out, p->zp_out, p->post_act, p->post_s_out, p->post_zp_out);
    return q;
}

int oracle_requant_tables(const oracle_conv_params *p, const int8_t *wt, const int32_t *bias,
                          int taps_per_o, float *mult, float *badd, int32_t *ibias)
{
    for (int o = 0; o < p->o; o++) {
        int qi = p->w_channels > 1 ? o : 0;
        double sw = p->s_w[qi];
        /* bias scale defaults to s_in * s_w as a FLOAT product: it lives in the float field
         * bias->qinfo[].scale (tests/utils/test_utils.c:660 convert_f32_bias) */
        const float sb_default = p->s_in * p->s_w[qi];
        double sb = p->s_b ? (double)p->s_b[qi] : (double)sb_default;
        int64_t wsum = 0;
        for (int t = 0; t < taps_per_o; t++) wsum += wt[(int64_t)o * taps_per_o + t];
        int64_t b = bias ? bias[o] : 0;
        /* a folded bias is un-folded exactly (reference does it in f32: convolution.c:375-395) */
        const int zw = p->zp_w ? p->zp_w[qi] : 0;
        /* the reference un-folds with the dequantised kernel, i.e. with sum (w - zp_w) (convolution.c:375-395) */
        if (p->fuse_zp2bias) b += (int64_t)p->zp_in * (wsum - (int64_t)taps_per_o * zw);
        mult[o] = (float)((double)p->s_in * sw / (double)p->s_out);
        badd[o] = (float)((double)b * sb / (double)p->s_out);
        ibias[o] = (int32_t)(-(int64_t)p->zp_in * wsum);
        int64_t wabs = 0;
        for (int t = 0; t < taps_per_o; t++) wabs += llabs((long long)wt[(int64_t)o * taps_per_o + t] - zw);
        const int xmax = (127 - p->zp_in) > (p->zp_in + 128) ? (127 - p->zp_in) : (p->zp_in + 128);
        double bound = (double)wabs * xmax * fabs(mult[o]) + fabs(badd[o]);
        if (!(bound < 8388608.0)) return -1; /* product refuses these too (b200_opt/quant.c) */
    }
    return 0;
}

int oracle_conv2d_i8(const oracle_conv_params *p, const int8_t *in, const int8_t *wt,
                     const int32_t *bias, int8_t *out)
{
    const int cg = p->c / p->group, og = p->o / p->group;
    const int taps = cg * p->kh * p->kw;
    float *mult = malloc(sizeof(float) * p->o), *badd = malloc(sizeof(float) * p->o);
    int32_t *ibias = malloc(sizeof(int32_t) * p->o);
    if (oracle_requant_tables(p, wt, bias, taps, mult, badd, ibias)) return -1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int o = 0; o < p->o; o++) {
            const int g = o / og;
            for (int oy = 0; oy < p->oh; oy++) {
                for (int ox = 0; ox < p->ow; ox++) {
                    int32_t acc = 0, xsum = 0;
                    for (int ci = 0; ci < cg; ci++) {
                        const int c = g * cg + ci;
                        for (int ky = 0; ky < p->kh; ky++) {
                            for (int kx = 0; kx < p->kw; kx++) {
                                int iy = oy * p->stride_h - p->pad_top + ky * p->dil_h;
                                int ix = ox * p->stride_w - p->pad_left + kx * p->dil_w;
                                int x = p->zp_in;
                                if (iy >= 0 && iy < p->h && ix >= 0 && ix < p->w)
                                    x = in[(((int64_t)b * p->c + c) * p->h + iy) * p->w + ix];
                                acc += x * wt[(((int64_t)o * cg + ci) * p->kh + ky) * p->kw + kx];
                                xsum += x;
                            }
                        }
                    }
                    acc += ibias[o];
                    if (p->zp_w) acc -= p->zp_w[p->w_channels > 1 ? o : 0] * (xsum - p->zp_in * taps);
                    out[(((int64_t)b * p->o + o) * p->oh + oy) * p->ow + ox] =
                        (int8_t)epilogue_i8(acc, mult[o], badd[o], p);
                }
            }
        }
    }
    free(mult), free(badd), free(ibias);
    return 0;
}

int oracle_dwconv2d_i8(const oracle_conv_params *p, const int8_t *in, const int8_t *wt,
                       const int32_t *bias, int8_t *out)
{
    const int dm = p->o / p->c; /* depth multiplier, convolution.c:229 */
    const int taps = p->kh * p->kw;
    float *mult = malloc(sizeof(float) * p->o), *badd = malloc(sizeof(float) * p->o);
    int32_t *ibias = malloc(sizeof(int32_t) * p->o);
    if (oracle_requant_tables(p, wt, bias, taps, mult, badd, ibias)) return -1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int o = 0; o < p->o; o++) {
            const int c = o / dm;
            for (int oy = 0; oy < p->oh; oy++) {
                for (int ox = 0; ox < p->ow; ox++) {
                    int32_t acc = 0, xsum = 0;
                    for (int ky = 0; ky < p->kh; ky++) {
                        for (int kx = 0; kx < p->kw; kx++) {
                            int iy = oy * p->stride_h - p->pad_top + ky * p->dil_h;
                            int ix = ox * p->stride_w - p->pad_left + kx * p->dil_w;
                            int x = p->zp_in;
                            if (iy >= 0 && iy < p->h && ix >= 0 && ix < p->w)
                                x = in[(((int64_t)b * p->c + c) * p->h + iy) * p->w + ix];
                            acc += x * wt[((int64_t)o * p->kh + ky) * p->kw + kx];
                            xsum += x;
                        }
                    }
                    acc += ibias[o];
                    if (p->zp_w) acc -= p->zp_w[p->w_channels > 1 ? o : 0] * (xsum - p->zp_in * taps);
                    out[(((int64_t)b * p->o + o) * p->oh + oy) * p->ow + ox] =
                        (int8_t)epilogue_i8(acc, mult[o], badd[o], p);
                }
            }
        }
    }
    free(mult), free(badd), free(ibias);
    return 0;
}

int oracle_fc_i8(const oracle_conv_params *p, const int8_t *in, const int8_t *wt,
                 const int32_t *bias, int8_t *out)
{
    float *mult = malloc(sizeof(float) * p->o), *badd = malloc(sizeof(float) * p->o);
    int32_t *ibias = malloc(sizeof(int32_t) * p->o);
    if (oracle_requant_tables(p, wt, bias, p->c, mult, badd, ibias)) return -1;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int o = 0; o < p->o; o++) {
            int32_t acc = 0, xsum = 0;
            for (int k = 0; k < p->c; k++) {
                acc += (int)in[(int64_t)b * p->c + k] * wt[(int64_t)o * p->c + k];
                xsum += in[(int64_t)b * p->c + k];
            }
            acc += ibias[o];
            if (p->zp_w) acc -= p->zp_w[p->w_channels > 1 ? o : 0] * (xsum - p->zp_in * p->c);
            out[(int64_t)b * p->o + o] = (int8_t)epilogue_i8(acc, mult[o], badd[o], p);
        }
    }
    free(mult), free(badd), free(ibias);
    return 0;
}

/* ---- float restatements (fp16 path) -------------------------------------------- */
static inline float act_f(float v, int act)
{
    if (act != ORACLE_ACT_NONE) v = v > 0 ? v : 0;
    if (act == ORACLE_ACT_RELU6) v = v < 6 ? v : 6;
    return v;
}

int oracle_conv2d_f32(const oracle_conv_params *p, const float *in, const float *wt,
                      const float *bias, float *out)
{
    const int cg = p->c / p->group, og = p->o / p->group;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int o = 0; o < p->o; o++) {
            const int g = o / og;
            for (int oy = 0; oy < p->oh; oy++) {
                for (int ox = 0; ox < p->ow; ox++) {
                    float acc = 0;
                    for (int ky = 0; ky < p->kh; ky++) {
                        for (int kx = 0; kx < p->kw; kx++) {
                            int iy = oy * p->stride_h - p->pad_top + ky * p->dil_h;
                            int ix = ox * p->stride_w - p->pad_left + kx * p->dil_w;
                            if (iy < 0 || iy >= p->h || ix < 0 || ix >= p->w) continue;
                            for (int ci = 0; ci < cg; ci++) {
                                const int c = g * cg + ci;
                                acc += in[(((int64_t)b * p->c + c) * p->h + iy) * p->w + ix] *
                                       wt[(((int64_t)o * cg + ci) * p->kh + ky) * p->kw + kx];
                            }
                        }
                    }
                    if (bias) acc += bias[o];
                    out[(((int64_t)b * p->o + o) * p->oh + oy) * p->ow + ox] = act_f(acc, p->act);
                }
            }
        }
    }
    return 0;
}

int oracle_dwconv2d_f32(const oracle_conv_params *p, const float *in, const float *wt,
                        const float *bias, float *out)
{
    const int dm = p->o / p->c;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int o = 0; o < p->o; o++) {
            const int c = o / dm;
            for (int oy = 0; oy < p->oh; oy++) {
                for (int ox = 0; ox < p->ow; ox++) {
                    float acc = 0;
                    for (int ky = 0; ky < p->kh; ky++) {
                        for (int kx = 0; kx < p->kw; kx++) {
                            int iy = oy * p->stride_h - p->pad_top + ky * p->dil_h;
                            int ix = ox * p->stride_w - p->pad_left + kx * p->dil_w;
                            if (iy < 0 || iy >= p->h || ix < 0 || ix >= p->w) continue;
                            acc += wt[((int64_t)o * p->kh + ky) * p->kw + kx] *
                                   in[(((int64_t)b * p->c + c) * p->h + iy) * p->w + ix];
                        }
                    }
                    if (bias) acc += bias[o];
                    out[(((int64_t)b * p->o + o) * p->oh + oy) * p->ow + ox] = act_f(acc, p->act);
                }
            }
        }
    }
    return 0;
}

int oracle_fc_f32(const oracle_conv_params *p, const float *in, const float *wt,
                  const float *bias, float *out)
{
#pragma omp parallel for schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int o = 0; o < p->o; o++) {
            float acc = 0;
            for (int k = 0; k < p->c; k++)
                acc += in[(int64_t)b * p->c + k] * wt[(int64_t)o * p->c + k];
            if (bias) acc += bias[o];
            out[(int64_t)b * p->o + o] = act_f(acc, p->act);
        }
    }
    return 0;
}

/* ---- elementwise / pooling: the reference's float sequence, verbatim in meaning ---- */
void oracle_relu_i8(const int8_t *in, int8_t *out, int64_t count, int act, float s_in, int zp_in,
                    float s_out, int zp_out)
{
    for (int64_t i = 0; i < count; i++)
        out[i] = (int8_t)post_stage(in[i], s_in, zp_in, act, s_out, zp_out);
}

/* leaky relu / sigmoid / clip through shl_ref_siso_callback_base (source/reference/utils.c:609):
 * dequantise, the f32 loop of leaky_relu.c:31-34 / sigmoid.c:31-34 / clip.c:31-40, requantise */
void oracle_unary_i8(const int8_t *in, int8_t *out, int64_t count, int op, float p0, float p1, float s_in,
                     int zp_in, float s_out, int zp_out)
{
    for (int64_t i = 0; i < count; i++) {
        float val = dequant_i8(in[i], s_in, zp_in);
        float r;
        switch (op) {
            case ORACLE_UNARY_LEAKY_RELU:
                r = val > 0 ? val : val * p0;
                break;
            case ORACLE_UNARY_SIGMOID:
                r = 1.0f / (1.0f + exp(-val));
                break;
            case ORACLE_UNARY_SILU: /* silu.c:31 */
                r = val / (1.0f + exp(-val));
                break;
            case ORACLE_UNARY_ERF: /* erf.c:31 */
                r = erf(val);
                break;
            default: /* ORACLE_UNARY_CLIP */
                if (val < p0)
                    r = p0;
                else if (val > p1)
                    r = p1;
                else
                    r = val;
        }
        out[i] = quant_i8(r, s_out, zp_out);
    }
}

void oracle_add_i8(const int8_t *a, const int8_t *b, int8_t *out, int64_t count, float s_a,
                   int zp_a, float s_b, int zp_b, float s_out, int zp_out)
{
    for (int64_t i = 0; i < count; i++) {
        float r = dequant_i8(a[i], s_a, zp_a) + dequant_i8(b[i], s_b, zp_b);
        out[i] = quant_i8(r, s_out, zp_out);
    }
}

/* sub / mul: sub.c:21-33, mul.c:21-33 (one f32 op between the dequantised operands) */
void oracle_binary_i8(int op, const int8_t *a, const int8_t *b, int8_t *out, int64_t count, float s_a,
                      int zp_a, float s_b, int zp_b, float s_out, int zp_out)
{
    for (int64_t i = 0; i < count; i++) {
        float x = dequant_i8(a[i], s_a, zp_a), y = dequant_i8(b[i], s_b, zp_b);
        float r = op == 1 ? x - y : (op == 2 ? x * y : (op == 3 ? (x >= 0 ? x : x * y) /* prelu.c:44-48 */ : x + y));
        if (op == 4) r = x / y; /* div.c:22 */
        /* 0 / 0: float_to_int8_base (source/nn2/utils.c:550) falls through both comparisons and casts NaN to
         * int8_t -- 0 on the reference's x86 build (cvttss2si gives 0x80000000, the low byte is kept) */
        out[i] = r != r ? 0 : quant_i8(r, s_out, zp_out);
    }
}

/* concat: source/reference/concat.c:20-48 (outer x [input i: dim_i[axis] * inner] copies) between the
 * per-input dequantisation and the output quantisation of shl_ref_concat_quant :52-80 */
void oracle_concat_i8(int k, const int8_t *const *in, const int64_t *axis_dim, const float *s_in,
                      const int32_t *zp_in, int64_t outer, int64_t inner, float s_out, int zp_out, int8_t *out)
{
    for (int64_t o = 0; o < outer; o++)
        for (int i = 0; i < k; i++) {
            const int64_t run = axis_dim[i] * inner;
            const int8_t *src = in[i] + o * run;
            for (int64_t j = 0; j < run; j++) *out++ = quant_i8(dequant_i8(src[j], s_in[i], zp_in[i]), s_out, zp_out);
        }
}

void oracle_avgpool_i8(const oracle_pool_params *p, const int8_t *in, int8_t *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int c = 0; c < p->c; c++) {
            for (int oy = 0; oy < p->oh; oy++) {
                for (int ox = 0; ox < p->ow; ox++) {
                    int x0 = ox * p->stride_w - p->pad_left, y0 = oy * p->stride_h - p->pad_top;
                    int fx0 = x0 < 0 ? -x0 : 0, fy0 = y0 < 0 ? -y0 : 0;
                    int fx1 = p->kw < p->w - x0 ? p->kw : p->w - x0;
                    int fy1 = p->kh < p->h - y0 ? p->kh : p->h - y0;
                    float total = 0.f, cnt = 0;
                    for (int fy = fy0; fy < fy1; fy++)
                        for (int fx = fx0; fx < fx1; fx++) {
                            total += dequant_i8(
                                in[(((int64_t)b * p->c + c) * p->h + y0 + fy) * p->w + x0 + fx],
                                p->s_in, p->zp_in);
                            cnt++;
                        }
                    if (p->count_include_pad) cnt = (float)(p->kh * p->kw);
                    float avg = total / cnt;
                    out[(((int64_t)b * p->c + c) * p->oh + oy) * p->ow + ox] =
                        quant_i8(avg, p->s_out, p->zp_out);
                }
            }
        }
    }
}

void oracle_maxpool_i8(const oracle_pool_params *p, const int8_t *in, int8_t *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < p->n; b++) {
        for (int c = 0; c < p->c; c++) {
            for (int oy = 0; oy < p->oh; oy++) {
                for (int ox = 0; ox < p->ow; ox++) {
                    int x0 = ox * p->stride_w - p->pad_left, y0 = oy * p->stride_h - p->pad_top;
                    int fx0 = x0 < 0 ? -x0 : 0, fy0 = y0 < 0 ? -y0 : 0;
                    int fx1 = p->kw < p->w - x0 ? p->kw : p->w - x0;
                    int fy1 = p->kh < p->h - y0 ? p->kh : p->h - y0;
                    float m = -FLT_MAX;
                    for (int fy = fy0; fy < fy1; fy++)
                        for (int fx = fx0; fx < fx1; fx++)
                            m = (float)fmax(
                                m,
                                dequant_i8(in[(((int64_t)b * p->c + c) * p->h + y0 + fy) * p->w +
                                              x0 + fx],
                                           p->s_in, p->zp_in));
                    out[(((int64_t)b * p->c + c) * p->oh + oy) * p->ow + ox] =
                        quant_i8(m, p->s_out, p->zp_out);
                }
            }
        }
    }
}

void oracle_softmax_i8(const int8_t *in, int8_t *out, int rows, int c, float s_in, int zp_in,
                       float s_out, int zp_out)
{
    for (int r = 0; r < rows; r++) {
        const int8_t *x = in + (int64_t)r * c;
        float acc = 0.0f, mx = -FLT_MAX;
        for (int j = 0; j < c; j++) mx = (float)fmax(mx, dequant_i8(x[j], s_in, zp_in));
        for (int j = 0; j < c; j++) acc += exp(dequant_i8(x[j], s_in, zp_in) - mx);
        for (int j = 0; j < c; j++) {
            float v = exp(dequant_i8(x[j], s_in, zp_in) - mx) / acc;
            out[(int64_t)r * c + j] = quant_i8(v, s_out, zp_out);
        }
    }
}

/* ---- fp16 conversions exactly as the reference (nn2/utils.c:576-645) ---------------- */
uint16_t oracle_f32_to_f16(float value)
{
    if (value > 65519.0) return 0x7BFF;
    if (value < -65519.0) return 0xFBFF;
    union {
        uint32_t u;
        float f;
    } in, magic;
    const uint32_t f32inf = 255u << 23, f16inf = 31u << 23;
    magic.u = 15u << 23;
    in.f = value;
    uint32_t sign = in.u & 0x80000000u;
    in.u ^= sign;
    uint16_t out;
    if (in.u >= f32inf) {
        out = in.u > f32inf ? 0x7FFF : 0x7C00;
    } else {
        in.u &= ~0xFFFu;
        in.f *= magic.f;
        in.u -= ~0xFFFu;
        if (in.u > f16inf) in.u = f16inf;
        out = (uint16_t)(in.u >> 13);
    }
    return (uint16_t)(out | (sign >> 16));
}

float oracle_f16_to_f32(uint16_t value)
{
    union {
        uint32_t u;
        float f;
    } out, magic, inf;
    magic.u = (254u - 15u) << 23;
    inf.u = (127u + 16u) << 23;
    out.u = (uint32_t)(value & 0x7FFF) << 13;
    out.f *= magic.f;
    if (out.f >= inf.f) out.u |= 255u << 23;
    out.u |= (uint32_t)(value & 0x8000) << 16;
    return out.f;
}
