/*
 * ops.c -- operator callbacks of the b200 backend: the init / exec pairs the reference's
 * front ends call (source/nn2/convolution.c:26-86, depthwise_conv2d.c, fullyconnected.c,
 * relu.c, add.c, maxpool.c, averagepool.c, global_avgpool.c, softmax.c, reshape.c, flatten.c).
 *
 *   init : validate, pack weights + per-channel tables into the device weight arena, build the
 *          b200_op, bind it to the params struct, set cb->exec.  Host kernel / bias buffers are
 *          left untouched (the RVV back end rewrites them in place,
 *          source/thead_rvv/int8/convolution.c:172-190).
 *   exec : layer mode (CSINN_RM_LAYER).  Synchronous by contract -- the caller reads
 *          output->data right after return (source/nn2/convolution.c:79) -- so it stages
 *          H2D -> NCHW->pixel-major -> kernels -> pixel-major->NCHW -> D2H -> stream sync.
 *          Graph mode never goes through exec: graph.c plans the same b200_ops once and
 *          replays them as a CUDA graph.
 * There is no CPU path: every failure returns CSINN_FALSE / a negative status after logging
 * through shl_debug_error, and shl_b200_last_error() keeps the message.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200_internal.h"

/* ---- errors --------------------------------------------------------------------------- */
/* The reference's front ends discard what init / exec return (source/nn2/convolution.c:50-55,
 * 64-86 always answer CSINN_TRUE), so a status code alone would be silent.  Every failure is
 * therefore (1) written to stderr unconditionally, (2) kept for shl_b200_last_error(), (3)
 * counted (shl_b200_error_count()), and (4) fatal when SHL_B200_ABORT_ON_ERROR is set. */
static char g_err[640];
static int g_err_count;
void b200_fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    g_err_count++;
    fprintf(stderr, "[shl_b200 ERROR] %s\n", g_err);
    if (getenv("SHL_B200_ABORT_ON_ERROR")) abort();
}
const char *shl_b200_last_error(void) { return g_err; }
int shl_b200_error_count(void) { return g_err_count; }
void shl_b200_clear_error(void) { g_err[0] = 0; }

#define DEV_CHECK(expr)                                                        \
    do {                                                                       \
        int _rc = (expr);                                                      \
        if (_rc != B200_OK) {                                                  \
            b200_fail("%s -> %d: %s", #expr, _rc, b200_last_error());          \
            return CSINN_FALSE;                                                \
        }                                                                      \
    } while (0)

/* ---- context --------------------------------------------------------------------------- */
static int g_device = -1;
int shl_b200_set_device(int device)
{
    g_device = device;
    return CSINN_TRUE;
}
int b200_default_device(void)
{
    if (g_device >= 0) return g_device;
    const char *e = getenv("SHL_B200_DEVICE");
    if (!e) e = getenv("LOCAL_RANK");
    return e ? atoi(e) : 0;
}

int b200_ctx_init(b200_ctx *ctx, int device)
{
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    int rc = b200_set_device(device);
    if (rc != B200_OK) {
        b200_fail("cannot use CUDA device %d: %s", device, b200_last_error());
        return CSINN_FALSE;
    }
    DEV_CHECK(b200_stream_create(&ctx->stream));
    return CSINN_TRUE;
}

void b200_ctx_destroy(b200_ctx *ctx)
{
    if (ctx->stream) b200_stream_destroy(ctx->stream);
    if (ctx->copy_stream) b200_stream_destroy(ctx->copy_stream);
    /* arena chunks of the default context live for the process; a session's fixed arena is
     * released here */
    if (ctx->fixed_arena && ctx->wbase) b200_free(ctx->wbase);
    memset(ctx, 0, sizeof(*ctx));
}

static b200_ctx g_ctx;
static int g_ctx_ready;
b200_ctx *b200_ctx_default(void)
{
    if (!g_ctx_ready) {
        if (b200_ctx_init(&g_ctx, b200_default_device()) != CSINN_TRUE) return NULL;
        g_ctx_ready = 1;
    }
    b200_set_device(g_ctx.device);
    return &g_ctx;
}

b200_ctx *b200_ctx_of(struct csinn_session *sess)
{
    b200_option *opt = b200_option_of(sess);
    if (opt) {
        b200_set_device(opt->ctx.device);
        return &opt->ctx;
    }
    return b200_ctx_default();
}

void *b200_warena_put(b200_ctx *ctx, const void *src, size_t bytes)
{
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (!ctx->wbase || ctx->wused + need > ctx->wcap) {
        if (ctx->fixed_arena) {
            b200_fail("weight arena exhausted (%zu + %zu > %zu)", ctx->wused, need, ctx->wcap);
            return NULL;
        }
        size_t cap = (size_t)64 << 20;
        if (cap < need) cap = need;
        void *p = NULL;
        if (b200_malloc(&p, cap) != B200_OK) {
            b200_fail("weight arena allocation of %zu bytes failed: %s", cap, b200_last_error());
            return NULL;
        }
        ctx->wbase = p, ctx->wcap = cap, ctx->wused = 0;
    }
    void *dst = ctx->wbase + ctx->wused;
    ctx->wused += need;
    if (src && !ctx->skip_upload) {
        /* pageable source: the runtime stages it, the call returns once the copy is queued
         * from a private staging buffer -- sync so the caller may free `src` at once */
        if (b200_memcpy_h2d(dst, src, bytes, ctx->stream) != B200_OK ||
            b200_stream_sync(ctx->stream) != B200_OK) {
            b200_fail("weight upload failed: %s", b200_last_error());
            return NULL;
        }
    } else if (b200_memset(dst, 0, need, ctx->stream) != B200_OK) {
        b200_fail("weight arena clear failed: %s", b200_last_error());
        return NULL;
    }
    return dst;
}

/* ---- device tensors --------------------------------------------------------------------- */
size_t b200_dt_bytes(const b200_dt *t)
{
    if (t->is_nchw) return (size_t)t->n * t->c * t->h * t->w * t->eb;
    return (size_t)t->n * t->h * t->w * t->cp * t->eb;
}

int b200_dt_from_tensor(b200_dt *t, const struct csinn_tensor *src)
{
    memset(t, 0, sizeof(*t));
    if (src->dtype == CSINN_DTYPE_INT8)
        t->eb = 1;
    else if (src->dtype == CSINN_DTYPE_FLOAT16)
        t->eb = 2;
    else
        return 0;
    switch (src->dim_count) {
        case 4:
            t->n = src->dim[0], t->c = src->dim[1], t->h = src->dim[2], t->w = src->dim[3];
            break;
        case 3:
            t->n = src->dim[0], t->c = src->dim[1], t->h = 1, t->w = src->dim[2];
            break;
        case 2:
            t->n = src->dim[0], t->c = src->dim[1], t->h = 1, t->w = 1;
            break;
        case 1:
            t->n = 1, t->c = src->dim[0], t->h = 1, t->w = 1;
            break;
        default:
            return 0;
    }
    if (t->n <= 0 || t->c <= 0 || t->h <= 0 || t->w <= 0) return 0;
    t->cp = b200_round_channels(t->c, t->eb);
    return 1;
}

/* ---- params -> op registry (open addressing, grows; tombstones on release) --------------- */
#define REG_TOMB ((void *)1)
static struct reg_slot {
    void *key;
    b200_op *op;
} *g_reg;
static size_t g_reg_cap, g_reg_used;

static size_t reg_hash(void *p, size_t cap) { return (size_t)(((uintptr_t)p >> 4) * 2654435761u) & (cap - 1); }

static void reg_grow(void)
{
    const size_t ncap = g_reg_cap ? g_reg_cap * 2 : 1024;
    struct reg_slot *n = calloc(ncap, sizeof(*n));
    if (!n) return;
    for (size_t i = 0; i < g_reg_cap; i++) {
        if (!g_reg[i].key || g_reg[i].key == REG_TOMB) continue;
        size_t h = reg_hash(g_reg[i].key, ncap);
        while (n[h].key) h = (h + 1) & (ncap - 1);
        n[h] = g_reg[i];
    }
    size_t live = 0;
    for (size_t i = 0; i < ncap; i++) live += n[i].key != NULL;
    free(g_reg);
    g_reg = n, g_reg_cap = ncap, g_reg_used = live;
}

void b200_op_bind(void *params, b200_op *op)
{
    if ((g_reg_used + 1) * 2 > g_reg_cap) reg_grow();
    if (!g_reg) {
        b200_fail("out of host memory growing the operator registry");
        return;
    }
    size_t h = reg_hash(params, g_reg_cap), tomb = (size_t)-1;
    while (g_reg[h].key && g_reg[h].key != params) {
        if (g_reg[h].key == REG_TOMB && tomb == (size_t)-1) tomb = h;
        h = (h + 1) & (g_reg_cap - 1);
    }
    if (!g_reg[h].key) {
        if (tomb != (size_t)-1)
            h = tomb;
        else
            g_reg_used++;
    }
    if (g_reg[h].key == params && g_reg[h].op && g_reg[h].op != op) {
        /* re-init of the same params (a second session_setup, a repeated layer-mode csinn_*_init): the old
         * operator's staging buffers go, its constants stay in the arena they were bumped into */
        b200_op *old = g_reg[h].op;
        b200_set_device(old->ctx->device);
        for (int i = 0; i < 8; i++)
            if (old->stg[i]) b200_free(old->stg[i]);
        free(old);
    }
    g_reg[h].key = params;
    g_reg[h].op = op;
}

b200_op *b200_op_find(void *params)
{
    if (!g_reg) return NULL;
    size_t h = reg_hash(params, g_reg_cap);
    while (g_reg[h].key) {
        if (g_reg[h].key == params) return g_reg[h].op;
        h = (h + 1) & (g_reg_cap - 1);
    }
    return NULL;
}

/* The reference API has no per-operator deinit (its back ends leak kernel_tm, see the FIXME at
 * source/thead_rvv/int8/convolution_gemm_int8.c:50); this addition lets long-running hosts and
 * the test harness drop an operator's staging buffers. */
void shl_b200_op_release(void *params)
{
    if (!g_reg) return;
    size_t h = reg_hash(params, g_reg_cap);
    while (g_reg[h].key) {
        if (g_reg[h].key == params) {
            b200_op *op = g_reg[h].op;
            g_reg[h].key = REG_TOMB, g_reg[h].op = NULL;
            if (op) {
                b200_set_device(op->ctx->device);
                for (int i = 0; i < 8; i++)
                    if (op->stg[i]) b200_free(op->stg[i]);
                free(op);
            }
            return;
        }
        h = (h + 1) & (g_reg_cap - 1);
    }
}

static b200_op *op_new(struct csinn_params_base *base, int kind, int csinn_dtype, const char *kname)
{
    int dt;
    if (csinn_dtype == CSINN_DTYPE_INT8)
        dt = B200_I8;
    else if (csinn_dtype == CSINN_DTYPE_FLOAT16)
        dt = B200_F16;
    else {
        b200_fail("dtype %d not supported by the b200 backend (int8 and float16 only)", csinn_dtype);
        return NULL;
    }
    b200_ctx *ctx = b200_ctx_of(base->sess);
    if (!ctx) return NULL;
    b200_op *op = calloc(1, sizeof(*op));
    if (!op) return NULL;
    op->kind = kind, op->dtype = dt, op->eb = dt == B200_I8 ? 1 : 2;
    op->kname = kname, op->ctx = ctx;
    return op;
}

static void fill_epilogue(const b200_op *op, b200_epilogue *ep)
{
    memset(ep, 0, sizeof(*ep));
    ep->mult = op->d_mult, ep->badd = op->d_badd, ep->ibias = op->d_ibias;
    ep->post_lut = op->kind == B200_OPK_ACT ? NULL : op->d_lut;
    ep->zp_out = op->zp_out, ep->act = op->act, ep->q6 = op->q6;
}

/* ---- running one op on device tensors ------------------------------------------------------ */
/* a network's first layer: NCHW input with so few channels that im2col + GEMM loses to the
 * direct dp4a kernel (csrc/conv_direct.cu) */
static int conv_goes_direct(const b200_op *op, const b200_dt *in0)
{
    if (op->d_wzp) return 0; /* asymmetric weights need the row sums of the im2col matrix: im2col + GEMM */
    if (!(in0->is_nchw && op->group == 1 && op->kdim <= 160 &&
          ((op->dtype == B200_I8 && op->o <= 256) || (op->dtype == B200_F16 && op->o <= 64)) &&
          !getenv("SHL_B200_NO_DIRECT_CONV")))
        return 0;
    if (op->dtype == B200_I8) {
        /* the dp4a kernel keeps the weights, the per-channel tables and (generic variant) one tap vector per
         * thread in shared memory and refuses above 48 KB (csrc/conv_direct.cu, b200_conv2d_direct): such a
         * layer takes im2col + GEMM instead.  kname, scratch planning and run_conv all decide through here. */
        const size_t kwords = (size_t)(op->kdim + 3) / 4, o4 = (size_t)(op->o + 3) & ~(size_t)3;
        const int special = in0->c == 3 && op->kh == op->kw && (op->kh == 3 || op->kh == 7) && op->dh == 1 && op->dw == 1;
        const size_t smem = kwords * o4 * 4 + o4 * 12 + (special ? 0 : kwords * 128 * 4);
        if (smem > 48 * 1024) return 0;
    }
    return 1;
}

/* the 3-channel 3x3 / 7x7 stems run as an implicit GEMM on the tensor core (csrc/conv_stem_tc.cu);
 * same conditions as b200_conv_stem_tc_launch (the shim decides, this only names the kernel) */
static int conv_stem_on_tc(const b200_op *op, const b200_dt *in0)
{
    return in0->c == 3 && op->kh == op->kw && (op->kh == 3 || op->kh == 7) && op->dh == 1 && op->dw == 1 &&
           op->sh == op->sw && op->sw <= 2 && (op->kh == 3 || op->sw == 2) && op->o <= 64 && in0->w % 4 == 0 &&
           !getenv("SHL_B200_NO_STEM_TC");
}

/* k > 1 convolutions whose channel count is a multiple of 64 gather their A operand by TMA im2col loads
 * (csrc/gemm_tc.cu, IGEMM) instead of writing an im2col matrix to scratch */
static int conv_uses_igemm(const b200_op *op, const b200_dt *in0)
{
    return op->kind == B200_OPK_CONV && op->ig_ncls >= 1 && !op->direct && !op->d_wzp && !in0->is_nchw &&
           in0->h == op->ig_h && in0->w == op->ig_w && in0->c == op->cin && !getenv("SHL_B200_NO_IGEMM");
}

/* Depthwise convolutions with channels a multiple of 64 CAN run on the implicit-GEMM kernel against tap-diagonal
 * weights (SHL_B200_DW_IGEMM=1; bit-exact, in the parity suite).  Off by default: measured 2x slower than the dp4a
 * kernel (14 x 14 x 512 at batch 256: 53 vs 26 us) -- nine im2col-mode TMA loads of 128 pixels per tile cost ~4.5 cycles
 * per pixel request in the TMA unit, the same rate that bounds the dense implicit GEMM. */
static int dw_uses_igemm(const b200_op *op, const b200_dt *in0)
{
    if (op->kind != B200_OPK_DW || !op->d_w_diag || op->ig_ncls < 1 || in0->is_nchw || in0->h != op->ig_h ||
        in0->w != op->ig_w)
        return 0;
    if (getenv("SHL_B200_DW_GENERIC") || getenv("SHL_B200_DW_UMMA")) return 0; /* another kernel was asked for */
    const char *e = getenv("SHL_B200_DW_IGEMM");
    return e && atoi(e) != 0;
}

const char *b200_op_kname(const b200_op *op, const b200_dt *in0)
{
    if (conv_uses_igemm(op, in0)) return "b200_conv_igemm_tcgen05";
    if (dw_uses_igemm(op, in0)) return "b200_dwconv_igemm_tcgen05";
    if (op->kind == B200_OPK_CONV && conv_goes_direct(op, in0))
        return (op->dtype == B200_I8 && conv_stem_on_tc(op, in0)) ? "b200_conv2d_stem_tcgen05" : "b200_conv2d_direct";
    return op->kname;
}

/* scratch layout: [im2col matrix of one group][row sums, int32 per output pixel (asymmetric weights only)] */
static size_t im2col_bytes(const b200_op *op, const b200_dt *in0, const b200_dt *out)
{
    if (op->kind != B200_OPK_CONV || (op->direct && !in0->is_nchw)) return 0;
    if (conv_goes_direct(op, in0) || conv_uses_igemm(op, in0)) return 0;
    return ((size_t)out->n * out->h * out->w * op->ldk * op->eb + 255) & ~(size_t)255;
}

/* group convolutions whose outputs per group are not a multiple of 16 bytes (this includes depthwise convolutions with a
 * depth multiplier, which are group convolutions with one input channel per group): a group's GEMM cannot store into
 * its column window of the output directly (the TMA store wants 16-byte aligned rows), so it writes a staging tensor
 * [pixels][og rounded up] that a slice copy then places -- slow, exact, and only on this rare path */
static int group_needs_stage(const b200_op *op)
{
    return op->kind == B200_OPK_CONV && op->group > 1 && ((op->o / op->group) * op->eb) % 16 != 0;
}
static size_t group_stage_bytes(const b200_op *op, const b200_dt *out)
{
    if (!group_needs_stage(op)) return 0;
    const size_t ldo = (size_t)b200_round_channels(op->o / op->group, op->eb);
    return ((size_t)out->n * out->h * out->w * ldo * op->eb + 255) & ~(size_t)255;
}

static size_t matmul_scratch(const b200_op *op, size_t *off_b, size_t *off_o, size_t *off_rs);
size_t b200_op_scratch_bytes(const b200_op *op, const b200_dt *in0, const b200_dt *out)
{
    if (op->kind == B200_OPK_TENSOR && op->t_op == B200_T_MATMUL) return matmul_scratch(op, NULL, NULL, NULL);
    if (op->kind == B200_OPK_COPY && op->direct == 2) /* general reshape: the tensor in NCHW order in between */
        return ((size_t)in0->n * in0->c * in0->h * in0->w * in0->eb + 255) & ~(size_t)255;
    if (op->kind != B200_OPK_CONV && op->kind != B200_OPK_FC) return 0;
    size_t bytes = im2col_bytes(op, in0, out);
    if (op->d_wzp) bytes += (size_t)out->n * out->h * out->w * sizeof(int32_t);
    bytes += group_stage_bytes(op, out);
    return bytes;
}

static int run_conv(b200_op *op, const b200_dt *in, const b200_dt *out, void *scratch, void *stream)
{
    const int og = op->o / op->group, cg = op->cin / op->group;
    const int m = out->n * out->h * out->w;
    if (conv_goes_direct(op, in)) {
        b200_conv_direct_desc c;
        memset(&c, 0, sizeof(c));
        c.n = in->n, c.c = in->c, c.h = in->h, c.w = in->w;
        c.o = op->o, c.oh = out->h, c.ow = out->w, c.cp_out = out->cp;
        c.kh = op->kh, c.kw = op->kw, c.stride_h = op->sh, c.stride_w = op->sw;
        c.pad_top = op->pt, c.pad_left = op->pl, c.dil_h = op->dh, c.dil_w = op->dw;
        c.ldw = op->ldk * op->eb, c.in = in->d, c.wt = op->d_w, c.out = out->d, c.zp_in = op->zp_in;
        fill_epilogue(op, &c.ep);
        if (op->dtype == B200_F16)
            DEV_CHECK(b200_conv2d_direct_f16(&c, stream));
        else
            DEV_CHECK(b200_conv2d_direct(&c, stream));
        return CSINN_TRUE;
    }
    if (conv_uses_igemm(op, in)) {
        b200_conv_igemm_desc c;
        memset(&c, 0, sizeof(c));
        c.n = in->n, c.h = in->h, c.w = in->w, c.c = in->c, c.cp_in = in->cp;
        c.o = op->o, c.oh = out->h, c.ow = out->w, c.kh = op->kh, c.kw = op->kw;
        c.stride_h = op->sh, c.stride_w = op->sw, c.pad_top = op->pt, c.pad_left = op->pl, c.dil_h = op->dh, c.dil_w = op->dw;
        c.in = in->d, c.wt = op->d_w, c.ldw = op->ldk, c.out = out->d, c.ldo = out->cp;
        fill_epilogue(op, &c.ep);
        c.ncls = op->ig_ncls, c.seeds = op->d_ig_seeds, c.cls_map = op->d_ig_clsmap;
        DEV_CHECK(b200_conv_igemm(&c, stream));
        return CSINN_TRUE;
    }
    b200_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.dtype = op->dtype, g.m = m, g.k = op->kdim, g.ldw = op->ldk, g.ldo = out->cp;
    fill_epilogue(op, &g.ep);
    for (int grp = 0; grp < op->group; grp++) {
        if (op->direct && !in->is_nchw) {
            g.a = in->d, g.lda = in->cp;
        } else {
            b200_im2col_desc c;
            memset(&c, 0, sizeof(c));
            c.dtype = op->dtype, c.n = in->n, c.h = in->h, c.w = in->w, c.cp_in = in->cp;
            c.in_nchw = in->is_nchw, c.c_total = in->c, c.c_off = grp * cg, c.cg = cg;
            c.oh = out->h, c.ow = out->w, c.kh = op->kh, c.kw = op->kw;
            c.stride_h = op->sh, c.stride_w = op->sw, c.pad_top = op->pt, c.pad_left = op->pl;
            c.dil_h = op->dh, c.dil_w = op->dw, c.ldk = op->ldk, c.pad_value = op->zp_in;
            c.in = in->d, c.col = scratch;
            DEV_CHECK(b200_im2col(&c, stream));
            g.a = scratch, g.lda = op->ldk;
        }
        if (op->d_wzp) {
            int32_t *rs = (int32_t *)((uint8_t *)scratch + im2col_bytes(op, in, out));
            DEV_CHECK(b200_rowsum_i8(g.a, g.lda, m, op->kdim, op->zp_in, rs, stream));
            g.w_zp = op->d_wzp + grp * og, g.rowsum = rs;
        }
        g.n = og;
        g.w = (const uint8_t *)op->d_w + (size_t)grp * og * op->ldk * op->eb;
        g.out = (uint8_t *)out->d + (size_t)grp * og * op->eb;
        if (group_needs_stage(op)) {
            /* the group's columns go through a staging tensor and a slice copy (see group_needs_stage) */
            uint8_t *stage_buf = (uint8_t *)scratch + im2col_bytes(op, in, out) + (op->d_wzp ? (size_t)m * sizeof(int32_t) : 0);
            const int ldo = b200_round_channels(og, op->eb);
            g.out = stage_buf, g.ldo = ldo, g.out_cols = 0;
            g.ep.mult = op->d_mult ? op->d_mult + grp * og : NULL;
            g.ep.badd = op->d_badd ? op->d_badd + grp * og : NULL;
            g.ep.ibias = op->d_ibias ? op->d_ibias + grp * og : NULL;
            DEV_CHECK(b200_gemm(&g, stream));
            b200_concat_desc cd;
            memset(&cd, 0, sizeof(cd));
            cd.dtype = op->dtype, cd.n = out->n, cd.c = og, cd.h = out->h, cd.w = out->w, cd.cp_in = ldo;
            cd.on = out->n, cd.oc = out->c, cd.oh = out->h, cd.ow = out->w, cd.cp_out = out->cp;
            cd.axis = 1, cd.offset = grp * og, cd.in = stage_buf, cd.out = out->d, cd.lut = NULL, cd.extract = 0;
            DEV_CHECK(b200_concat_slice(&cd, stream));
            continue;
        }
        if (op->group > 1) {
            /* each group writes its own column window of the pixel-major output: tiles wider than the
             * window are clipped at its end (the GEMM's n-tile may be wider than og, e.g. og = 48 in a 64-column
             * tile), not at the row pitch -- beyond the window lie the next group's columns */
            g.ldo = out->cp;
            g.out_cols = og;
            g.ep.mult = op->d_mult ? op->d_mult + grp * og : NULL;
            g.ep.badd = op->d_badd ? op->d_badd + grp * og : NULL;
            g.ep.ibias = op->d_ibias ? op->d_ibias + grp * og : NULL;
        }
        DEV_CHECK(b200_gemm(&g, stream));
    }
    return CSINN_TRUE;
}

/* ---- transpose / gather / reduce_sum / layer_norm / rms_norm / matmul ---------------------------------------- */
static void view_of(b200_view *v, int rank, const int *dim, const b200_dt *dt)
{
    memset(v, 0, sizeof(*v));
    v->rank = rank;
    for (int i = 0; i < rank; i++) v->dim[i] = dim[i];
    v->cp = dt->cp;
}

/* scratch of a matmul: [A rows (M x ldk)][B rows (batches_b x J x ldk, when mat1 is an activation)][output rows (M x ldo)]
 * [row sums (M int32, int8 with a mat1 zero point)] */
static size_t matmul_scratch(const b200_op *op, size_t *off_b, size_t *off_o, size_t *off_rs)
{
    const size_t M = (size_t)op->mm_batches * op->mm_i;
    const size_t ldo = (size_t)b200_round_channels(op->mm_j, op->eb);
    size_t o = 0;
    o += (M * op->ldk * op->eb + 255) & ~(size_t)255;
    if (off_b) *off_b = o;
    if (!op->mm_const_b) o += ((size_t)op->mm_batches_b * op->mm_j * op->ldk * op->eb + 255) & ~(size_t)255;
    if (off_o) *off_o = o;
    o += (M * ldo * op->eb + 255) & ~(size_t)255;
    if (off_rs) *off_rs = o;
    if (op->d_wzp) o += M * sizeof(int32_t);
    return o;
}

static int run_tensor_op(b200_op *op, const b200_dt *in0, const b200_dt *in1, const b200_dt *out, void *scratch, void *stream)
{
    b200_view vi, vo;
    view_of(&vi, op->in_rank, op->in_dim, in0);
    view_of(&vo, op->out_rank, op->out_dim, out);
    /* the padding lanes of the output's channel pitch are not covered by the logical walks */
    DEV_CHECK(b200_memset(out->d, 0, b200_dt_bytes(out), stream));
    switch (op->t_op) {
        case B200_T_TRANSPOSE:
            DEV_CHECK(b200_permute(&vi, in0->d, &vo, out->d, op->t_perm, op->eb, op->d_lut, stream));
            return CSINN_TRUE;
        case B200_T_GATHER:
            DEV_CHECK(b200_gather(&vi, in0->d, &vo, out->d, op->t_axis, op->d_idx, op->n_idx, op->eb, op->d_lut, op->oob_q,
                                  stream));
            return CSINN_TRUE;
        case B200_T_REDUCE_SUM:
            DEV_CHECK(b200_reduce_sum(&vi, in0->d, &vo, out->d, op->t_axis, op->eb, op->s_in, op->zp_in, op->s_out,
                                      op->zp_out, stream));
            return CSINN_TRUE;
        case B200_T_LAYER_NORM:
        case B200_T_RMS_NORM:
            DEV_CHECK(b200_norm(op->t_op == B200_T_RMS_NORM, &vi, in0->d, out->d, op->t_axis, op->t_eps, op->d_gamma,
                                op->d_beta, op->eb, op->s_in, op->zp_in, op->s_out, op->zp_out, stream));
            return CSINN_TRUE;
        case B200_T_MATMUL: {
            /* rows of mat0 -> K-major A, the tcgen05 GEMM against the packed mat1 (constant: packed at init like
             * fullyconnected weights; activation: packed per run), rows back into the output tensor */
            size_t off_b = 0, off_o = 0, off_rs = 0;
            matmul_scratch(op, &off_b, &off_o, &off_rs);
            uint8_t *sc = scratch;
            if (!sc) {
                b200_fail("matmul: no scratch");
                return CSINN_FALSE;
            }
            const int M = op->mm_batches * op->mm_i;
            const int ldo = b200_round_channels(op->mm_j, op->eb);
            DEV_CHECK(b200_pack_rows(&vi, in0->d, op->mm_batches, op->mm_i, op->mm_k, op->mm_trans_a, sc, op->ldk, op->eb, stream));
            const void *w = op->d_w;
            if (!op->mm_const_b) {
                b200_view v1;
                if (!in1) {
                    b200_fail("matmul: second operand missing");
                    return CSINN_FALSE;
                }
                view_of(&v1, op->in1_rank, op->in1_dim, in1);
                /* mat1 is [.., K, J] (rows = j need a transposing walk) or, with trans_b, [.., J, K] */
                DEV_CHECK(b200_pack_rows(&v1, in1->d, op->mm_batches_b, op->mm_j, op->mm_k, !op->mm_trans_b, sc + off_b, op->ldk,
                                         op->eb, stream));
                w = sc + off_b;
            }
            b200_gemm_desc g;
            memset(&g, 0, sizeof(g));
            g.dtype = op->dtype, g.n = op->mm_j, g.k = op->mm_k, g.lda = op->ldk, g.ldw = op->ldk, g.ldo = ldo;
            g.w_dynamic = !op->mm_const_b; /* packed by the kernel just before: no weight prefetch under dependent launch */
            fill_epilogue(op, &g.ep);
            const int per_batch = !op->mm_const_b && op->mm_batches_b > 1;
            const int calls = per_batch ? op->mm_batches : 1;
            for (int b = 0; b < calls; b++) {
                g.m = per_batch ? op->mm_i : M;
                g.a = sc + (size_t)b * op->mm_i * op->ldk * op->eb;
                g.w = (const uint8_t *)w + (per_batch ? (size_t)b * op->mm_j * op->ldk * op->eb : 0);
                g.out = sc + off_o + (size_t)b * op->mm_i * ldo * op->eb;
                if (op->d_wzp) {
                    int32_t *rs = (int32_t *)(sc + off_rs) + (size_t)b * op->mm_i;
                    DEV_CHECK(b200_rowsum_i8(g.a, g.lda, g.m, op->mm_k, op->zp_in, rs, stream));
                    g.w_zp = op->d_wzp, g.rowsum = rs;
                }
                DEV_CHECK(b200_gemm(&g, stream));
            }
            DEV_CHECK(b200_unpack_rows(&vo, out->d, M, op->mm_j, sc + off_o, ldo, op->eb, stream));
            return CSINN_TRUE;
        }
    }
    b200_fail("unknown tensor op %d", op->t_op);
    return CSINN_FALSE;
}

int b200_op_run(b200_op *op, int part, const b200_dt *in0, const b200_dt *in1, const b200_dt *out,
                void *scratch, void *stream)
{
    switch (op->kind) {
        case B200_OPK_SPLIT: {
            if (part < 0 || part >= op->cat_n) {
                b200_fail("split: output %d of %d", part, op->cat_n);
                return CSINN_FALSE;
            }
            b200_concat_desc c;
            memset(&c, 0, sizeof(c));
            c.dtype = op->dtype, c.extract = 1;
            c.n = out->n, c.c = out->c, c.h = out->h, c.w = out->w, c.cp_out = out->cp;
            c.on = in0->n, c.oc = in0->c, c.oh = in0->h, c.ow = in0->w, c.cp_in = in0->cp;
            c.axis = op->cat_axis, c.offset = op->cat_off[part];
            c.in = in0->d, c.out = out->d, c.lut = op->cat_lut[part];
            DEV_CHECK(b200_concat_slice(&c, stream));
            return CSINN_TRUE;
        }
        case B200_OPK_CONCAT: {
            if (part < 0 || part >= op->cat_n) {
                b200_fail("concat: input %d of %d", part, op->cat_n);
                return CSINN_FALSE;
            }
            b200_concat_desc c;
            memset(&c, 0, sizeof(c));
            c.dtype = op->dtype;
            c.n = in0->n, c.c = in0->c, c.h = in0->h, c.w = in0->w, c.cp_in = in0->cp;
            c.on = out->n, c.oc = out->c, c.oh = out->h, c.ow = out->w, c.cp_out = out->cp;
            c.axis = op->cat_axis, c.offset = op->cat_off[part];
            c.in = in0->d, c.out = out->d, c.lut = op->cat_lut[part];
            DEV_CHECK(b200_concat_slice(&c, stream));
            return CSINN_TRUE;
        }
        case B200_OPK_CONV:
            return run_conv(op, in0, out, scratch, stream);
        case B200_OPK_FC: {
            b200_gemm_desc g;
            memset(&g, 0, sizeof(g));
            g.dtype = op->dtype, g.m = in0->n * in0->h * in0->w, g.n = op->o, g.k = op->kdim;
            g.a = in0->d, g.lda = in0->cp, g.w = op->d_w, g.ldw = op->ldk;
            g.out = out->d, g.ldo = out->cp;
            fill_epilogue(op, &g.ep);
            if (op->d_wzp) {
                DEV_CHECK(b200_rowsum_i8(g.a, g.lda, g.m, op->kdim, op->zp_in, (int32_t *)scratch, stream));
                g.w_zp = op->d_wzp, g.rowsum = (const int32_t *)scratch;
            }
            DEV_CHECK(b200_gemm(&g, stream));
            return CSINN_TRUE;
        }
        case B200_OPK_DW: {
            if (dw_uses_igemm(op, in0)) {
                b200_conv_igemm_desc c;
                memset(&c, 0, sizeof(c));
                c.n = in0->n, c.h = in0->h, c.w = in0->w, c.c = in0->c, c.cp_in = in0->cp;
                c.o = op->o, c.oh = out->h, c.ow = out->w, c.kh = op->kh, c.kw = op->kw;
                c.stride_h = op->sh, c.stride_w = op->sw, c.pad_top = op->pt, c.pad_left = op->pl, c.dil_h = op->dh, c.dil_w = op->dw;
                c.in = in0->d, c.wt = op->d_w_diag, c.ldw = op->ldk_diag, c.out = out->d, c.ldo = out->cp;
                fill_epilogue(op, &c.ep);
                c.ncls = op->ig_ncls, c.seeds = op->d_ig_seeds, c.cls_map = op->d_ig_clsmap, c.dw_slab = 1;
                DEV_CHECK(b200_conv_igemm(&c, stream));
                return CSINN_TRUE;
            }
            b200_dwconv_desc d;
            memset(&d, 0, sizeof(d));
            d.dtype = op->dtype, d.n = in0->n, d.c = in0->c, d.cp = in0->cp;
            d.h = in0->h, d.w = in0->w, d.oh = out->h, d.ow = out->w;
            d.kh = op->kh, d.kw = op->kw, d.stride_h = op->sh, d.stride_w = op->sw;
            d.pad_top = op->pt, d.pad_left = op->pl, d.dil_h = op->dh, d.dil_w = op->dw;
            d.in = in0->d, d.wt = op->d_w, d.wt_row3 = op->d_w2, d.out = out->d, d.zp_in = op->zp_in;
            d.w_zp = op->d_wzp;
            fill_epilogue(op, &d.ep);
            DEV_CHECK(b200_dwconv2d(&d, stream));
            return CSINN_TRUE;
        }
        case B200_OPK_ACT:
            if (op->dtype == B200_I8)
                DEV_CHECK(b200_lut_i8(in0->d, out->d, b200_dt_bytes(out), op->d_lut, stream));
            else if (op->act == B200_ACT_RELU || op->act == B200_ACT_RELU6)
                DEV_CHECK(b200_relu_f16(in0->d, out->d, b200_dt_bytes(out) / 2, op->act, stream));
            else
                DEV_CHECK(b200_unary_f16(in0->d, out->d, b200_dt_bytes(out) / 2, op->act, op->act_p0, op->act_p1, stream));
            return CSINN_TRUE;
        case B200_OPK_ADD:
            if (!op->d_const && !in1) {
                b200_fail("%s: second operand missing", op->kname);
                return CSINN_FALSE;
            }
            if (op->d_const && in0->cp != op->const_count) {
                b200_fail("%s: constant operand packed for %d channels, the activation has %d", op->kname, op->const_count, in0->cp);
                return CSINN_FALSE;
            }
            if (op->bcast_nc) { /* second activation: one pixel of channels per image */
                if (in1->cp != in0->cp || in1->h * in1->w != 1) {
                    b200_fail("%s: broadcast operand is not [N or 1, C, 1, 1]", op->kname);
                    return CSINN_FALSE;
                }
                DEV_CHECK(b200_binary_bcast_nc(op->binop, op->dtype, in0->d, in1->d, (size_t)in0->cp,
                                               (size_t)in0->h * in0->w * in0->cp, out->d,
                                               b200_dt_bytes(out) / op->eb, op->s_in, op->zp_in, op->s_in1, op->zp_in1, op->s_out,
                                               op->zp_out, op->d_lut, op->act, stream));
                return CSINN_TRUE;
            }
            DEV_CHECK(b200_binary_bcast(op->binop, op->dtype, in0->d, op->d_const ? op->d_const : in1->d,
                                        op->d_const ? (size_t)op->const_count : 0, out->d, b200_dt_bytes(out) / op->eb,
                                        op->s_in, op->zp_in, op->s_in1, op->zp_in1, op->s_out, op->zp_out, op->d_lut,
                                        op->act, stream));
            return CSINN_TRUE;
        case B200_OPK_POOL: {
            b200_pool_desc p;
            memset(&p, 0, sizeof(p));
            p.dtype = op->dtype, p.n = in0->n, p.c = in0->c, p.cp = in0->cp, p.h = in0->h, p.w = in0->w;
            p.oh = out->h, p.ow = out->w;
            p.kh = op->pool_global ? in0->h : op->kh, p.kw = op->pool_global ? in0->w : op->kw;
            p.stride_h = op->pool_global ? 1 : op->sh, p.stride_w = op->pool_global ? 1 : op->sw;
            p.pad_top = op->pool_global ? 0 : op->pt, p.pad_left = op->pool_global ? 0 : op->pl;
            p.is_avg = op->pool_avg, p.count_include_pad = op->count_include_pad;
            p.s_in = op->s_in, p.zp_in = op->zp_in, p.s_out = op->s_out, p.zp_out = op->zp_out;
            p.in = in0->d, p.out = out->d;
            DEV_CHECK(b200_pool2d(&p, stream));
            return CSINN_TRUE;
        }
        case B200_OPK_SOFTMAX:
            DEV_CHECK(b200_softmax(op->dtype, in0->d, out->d, in0->n * in0->h * in0->w, in0->c, in0->cp, out->cp,
                                   op->s_in, op->zp_in, op->s_out, op->zp_out, stream));
            return CSINN_TRUE;
        case B200_OPK_TENSOR:
            return run_tensor_op(op, in0, in1, out, scratch, stream);
        case B200_OPK_COPY:
            if (op->direct == 1) { /* flatten of an N x C x H x W tensor: back to NCHW order = the flattened rows */
                DEV_CHECK(b200_nhwc_to_nchw(in0->d, out->d, in0->n, in0->c, in0->h, in0->w, in0->cp, in0->eb, stream));
                return CSINN_TRUE;
            }
            if (op->direct == 2) { /* any other reshape: pixel-major -> NCHW order (= the API's row-major bytes, which a
                                    * reshape keeps) -> pixel-major of the new shape */
                if (!scratch) {
                    b200_fail("reshape: no scratch buffer planned");
                    return CSINN_FALSE;
                }
                DEV_CHECK(b200_nhwc_to_nchw(in0->d, scratch, in0->n, in0->c, in0->h, in0->w, in0->cp, in0->eb, stream));
                DEV_CHECK(b200_nchw_to_nhwc(scratch, out->d, out->n, out->c, out->h, out->w, out->cp, out->eb, 0, stream));
                return CSINN_TRUE;
            }
            if (b200_dt_bytes(in0) != b200_dt_bytes(out)) {
                b200_fail("reshape changes the device footprint (%zu -> %zu bytes)", b200_dt_bytes(in0),
                          b200_dt_bytes(out));
                return CSINN_FALSE;
            }
            DEV_CHECK(b200_memcpy_d2d(out->d, in0->d, b200_dt_bytes(out), stream));
            return CSINN_TRUE;
    }
    b200_fail("unknown op kind %d", op->kind);
    return CSINN_FALSE;
}

/* ---- depthwise 3x3 -> pointwise 1x1 as one kernel ---------------------------- */
static void dwpw_fill(const b200_op *dw, const b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out,
                      b200_dwpw_desc *f)
{
    memset(f, 0, sizeof(*f));
    b200_dwconv_desc *d = &f->dw;
    d->dtype = dw->dtype, d->n = in->n, d->c = in->c, d->cp = in->cp;
    d->h = in->h, d->w = in->w, d->oh = mid->h, d->ow = mid->w;
    d->kh = dw->kh, d->kw = dw->kw, d->stride_h = dw->sh, d->stride_w = dw->sw;
    d->pad_top = dw->pt, d->pad_left = dw->pl, d->dil_h = dw->dh, d->dil_w = dw->dw;
    d->in = in->d, d->wt = dw->d_w, d->wt_row3 = dw->d_w2, d->out = NULL, d->zp_in = dw->zp_in;
    fill_epilogue(dw, &d->ep);
    f->o = pw->o, f->w = pw->d_w, f->ldw = pw->ldk, f->out = out->d, f->ldo = out->cp;
    fill_epilogue(pw, &f->ep);
}

/* `mid` = the depthwise output as the graph declares it (shape only: it is never materialised).
 * SHL_B200_DWPW=0 never fuses, =1 fuses every pair the kernel covers; unset, the planner times both ways on
 * scratch buffers of the real sizes and keeps the faster (b200_dwpw_prefers_fusion) -- the same policy the
 * session applies to programmatic dependent launch. */
int b200_dwpw_can_fuse(const b200_op *dw, const b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out)
{
    /* SHL_B200_DWPW: unset / 0 = two kernels (the default since the fused kernel measured slower on every MobileNetV1
     * pair at batch 256), 1 = fuse every covered pair, "auto" = time both per pair at session_setup */
    const char *mode = getenv("SHL_B200_DWPW");
    if (getenv("SHL_B200_NO_DWPW") || !mode || (strcmp(mode, "auto") != 0 && atoi(mode) == 0)) return 0;
    if (dw->d_wzp || pw->d_wzp) return 0; /* asymmetric weights: two kernels (generic depthwise, row-sum GEMM) */
    if (dw->kind != B200_OPK_DW || pw->kind != B200_OPK_CONV || !pw->direct || dw->dtype != B200_I8 ||
        pw->dtype != B200_I8 || in->is_nchw || pw->kdim != mid->c || mid->c != in->c || mid->n != out->n ||
        mid->h != out->h || mid->w != out->w)
        return 0;
    b200_dwpw_desc f;
    dwpw_fill(dw, pw, in, mid, out, &f);
    f.dw.in = f.out = (void *)16; /* planning time: the arena is not allocated yet, only the shapes matter */
    return b200_dwpw_supported(&f);
}

/* device time of `reps` runs of the pair, fused (one kernel) or not (two), on the given buffers; < 0 on error */
static float dwpw_time(b200_op *dw, b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out, int fused,
                       int reps, void *stream, void *e0, void *e1)
{
    float ms = -1.f;
    for (int r = -1; r < reps; r++) { /* r = -1: warm-up (function attributes, caches) */
        if (r == 0 && b200_event_record(e0, stream) != B200_OK) return -1.f;
        if (fused) {
            if (b200_dwpw_run(dw, pw, in, mid, out, stream) != CSINN_TRUE) return -1.f;
        } else if (b200_op_run(dw, 0, in, NULL, mid, NULL, stream) != CSINN_TRUE ||
                   b200_op_run(pw, 0, mid, NULL, out, NULL, stream) != CSINN_TRUE)
            return -1.f;
    }
    if (b200_event_record(e1, stream) != B200_OK || b200_stream_sync(stream) != B200_OK ||
        b200_event_elapsed_ms(e0, e1, &ms) != B200_OK)
        return -1.f;
    return ms / reps;
}

/* 1: the fused kernel is at least as fast as depthwise + GEMM for this pair on this device (measured), 0: not */
int b200_dwpw_prefers_fusion(b200_op *dw, b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out)
{
    const char *mode = getenv("SHL_B200_DWPW");
    if (!mode) return 0;
    if (strcmp(mode, "auto") != 0) return atoi(mode) != 0;
    b200_dt ti = *in, tm = *mid, to = *out;
    void *e0 = NULL, *e1 = NULL;
    void *stream = dw->ctx->stream;
    int prefer = 0;
    ti.d = tm.d = to.d = NULL;
    const size_t bi = b200_dt_bytes(in);
    uint8_t *host = malloc(bi);
    if (!host) return 0;
    uint32_t lcg = 12345u; /* bytes that look like activations: a table epilogue's bank conflicts depend on the data */
    for (size_t i = 0; i < bi; i++) host[i] = (uint8_t)((lcg = lcg * 1664525u + 1013904223u) >> 24);
    if (b200_malloc(&ti.d, bi) == B200_OK && b200_malloc(&tm.d, b200_dt_bytes(mid)) == B200_OK &&
        b200_malloc(&to.d, b200_dt_bytes(out)) == B200_OK && b200_event_create(&e0) == B200_OK &&
        b200_event_create(&e1) == B200_OK && b200_memcpy_h2d(ti.d, host, bi, stream) == B200_OK &&
        b200_stream_sync(stream) == B200_OK) {
        const float t_fused = dwpw_time(dw, pw, &ti, &tm, &to, 1, 3, stream, e0, e1);
        const float t_plain = dwpw_time(dw, pw, &ti, &tm, &to, 0, 3, stream, e0, e1);
        prefer = t_fused > 0.f && t_plain > 0.f && t_fused <= t_plain;
        if (getenv("SHL_B200_DWPW_VERBOSE"))
            fprintf(stderr, "[shl_b200] dw3x3 -> 1x1 pair c=%d o=%d %dx%d: fused %.1f us, two kernels %.1f us -> %s\n", in->c,
                    out->c, out->h, out->w, t_fused * 1e3f, t_plain * 1e3f, prefer ? "fused" : "two kernels");
    }
    free(host);
    if (e0) b200_event_destroy(e0);
    if (e1) b200_event_destroy(e1);
    if (ti.d) b200_free(ti.d);
    if (tm.d) b200_free(tm.d);
    if (to.d) b200_free(to.d);
    return prefer;
}

int b200_dwpw_run(b200_op *dw, b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out, void *stream)
{
    b200_dwpw_desc f;
    dwpw_fill(dw, pw, in, mid, out, &f);
    DEV_CHECK(b200_dwpw_fused(&f, stream));
    return CSINN_TRUE;
}

int b200_op_can_fuse_act(const b200_op *op)
{
    return (op->kind == B200_OPK_CONV || op->kind == B200_OPK_DW || op->kind == B200_OPK_FC ||
            op->kind == B200_OPK_ADD) &&
           op->d_lut == NULL;
}

static int8_t *upload_lut(b200_ctx *ctx, int act, float p0, float p1, float s_in, int zp_in, float s_out, int zp_out)
{
    int8_t lut[256];
    b200_build_unary_lut(lut, act, p0, p1, s_in, zp_in, s_out, zp_out);
    return b200_warena_put(ctx, lut, 256);
}

int b200_op_fuse_act(b200_op *op, int act, float p0, float p1, const struct csinn_tensor *act_in,
                     const struct csinn_tensor *act_out)
{
    if (op->dtype == B200_F16) {
        /* relu on top of an already fused relu6 etc. is not produced by the planner; the fp16
         * epilogues know relu / relu6 only (leaky relu, sigmoid, clip stay separate steps) */
        if (op->act != B200_ACT_NONE || (act != B200_ACT_RELU && act != B200_ACT_RELU6)) return CSINN_FALSE;
        op->act = act;
        return CSINN_TRUE;
    }
    /* a relu / relu6 node that keeps its producer's qinfo is exactly the in-domain clamp the
     * epilogue already knows (max(q, zp), min(q, q6)): compare the tables and skip the lookup */
    if (op->act == B200_ACT_NONE && op->kind != B200_OPK_ADD && (act == B200_ACT_RELU || act == B200_ACT_RELU6)) {
        int8_t lut[256];
        b200_build_requant_lut(lut, act, act_in->qinfo->scale, act_in->qinfo->zero_point, act_out->qinfo->scale,
                               act_out->qinfo->zero_point);
        int same = 1;
        for (int q = -128; q < 128 && same; q++) {
            int v = q > op->zp_out ? q : op->zp_out;
            if (act == B200_ACT_RELU6 && v > op->q6) v = op->q6;
            same = lut[q + 128] == v;
        }
        if (same) {
            op->act = act;
            return CSINN_TRUE;
        }
    }
    op->d_lut = upload_lut(op->ctx, act, p0, p1, act_in->qinfo->scale, act_in->qinfo->zero_point,
                           act_out->qinfo->scale, act_out->qinfo->zero_point);
    return op->d_lut ? CSINN_TRUE : CSINN_FALSE;
}

/* ---- layer-mode execution ------------------------------------------------------------------ */
static void *stage(b200_op *op, int slot, size_t bytes)
{
    if (op->stg_bytes[slot] < bytes) {
        if (op->stg[slot]) b200_free(op->stg[slot]);
        op->stg[slot] = NULL, op->stg_bytes[slot] = 0;
        if (b200_malloc(&op->stg[slot], bytes) != B200_OK) {
            b200_fail("staging allocation of %zu bytes failed: %s", bytes, b200_last_error());
            return NULL;
        }
        op->stg_bytes[slot] = bytes;
    }
    return op->stg[slot];
}

static int upload_nchw(b200_op *op, int slot, const struct csinn_tensor *t, b200_dt *dt, void *stream)
{
    if (!b200_dt_from_tensor(dt, t)) {
        b200_fail("unsupported tensor (dtype %d, rank %d)", t->dtype, t->dim_count);
        return CSINN_FALSE;
    }
    if (!t->data) {
        b200_fail("tensor '%s' has no host data", t->name ? t->name : "?");
        return CSINN_FALSE;
    }
    const size_t raw = (size_t)dt->n * dt->c * dt->h * dt->w * dt->eb;
    void *d_raw = stage(op, slot, raw);
    void *d_pm = stage(op, slot + 1, b200_dt_bytes(dt));
    if (!d_raw || !d_pm) return CSINN_FALSE;
    DEV_CHECK(b200_memcpy_h2d(d_raw, t->data, raw, stream));
    DEV_CHECK(b200_nchw_to_nhwc(d_raw, d_pm, dt->n, dt->c, dt->h, dt->w, dt->cp, dt->eb, 0, stream));
    dt->d = d_pm;
    return CSINN_TRUE;
}

static int layer_exec(void *params, struct csinn_tensor *in0, struct csinn_tensor *in1,
                      struct csinn_tensor *output)
{
    b200_op *op = b200_op_find(params);
    if (!op) {
        b200_fail("exec before init: no b200 operator bound to these params");
        return CSINN_FALSE;
    }
    b200_set_device(op->ctx->device);
    void *stream = op->ctx->stream;
    b200_dt d0, d1, dout;
    memset(&d1, 0, sizeof(d1));
    if (upload_nchw(op, 0, in0, &d0, stream) != CSINN_TRUE) return CSINN_FALSE;
    if (in1 && upload_nchw(op, 2, in1, &d1, stream) != CSINN_TRUE) return CSINN_FALSE;
    if (!b200_dt_from_tensor(&dout, output) || !output->data) {
        b200_fail("unsupported or unallocated output tensor");
        return CSINN_FALSE;
    }
    const size_t raw = (size_t)dout.n * dout.c * dout.h * dout.w * dout.eb;
    dout.d = stage(op, 4, b200_dt_bytes(&dout));
    void *d_raw = stage(op, 5, raw);
    if (!dout.d || !d_raw) return CSINN_FALSE;
    void *scratch = NULL;
    const size_t sb = b200_op_scratch_bytes(op, &d0, &dout);
    if (sb && !(scratch = stage(op, 6, sb))) return CSINN_FALSE;
    if (b200_op_run(op, 0, &d0, in1 ? &d1 : NULL, &dout, scratch, stream) != CSINN_TRUE) return CSINN_FALSE;
    DEV_CHECK(b200_nhwc_to_nchw(dout.d, d_raw, dout.n, dout.c, dout.h, dout.w, dout.cp, dout.eb, stream));
    DEV_CHECK(b200_memcpy_d2h(output->data, d_raw, raw, stream));
    DEV_CHECK(b200_stream_sync(stream));
    return CSINN_TRUE;
}

/* ---- conv2d ---------------------------------------------------------------------------------- */
static int conv_act_of(struct csinn_params_base *base, int op_relu, int op_relu6)
{
    /* the front end does not pass the op enum to init; CONV2D_RELU / _RELU6 are registered with
     * their own init wrappers below */
    (void)base;
    (void)op_relu;
    (void)op_relu6;
    return B200_ACT_NONE;
}

static int conv_init_impl(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_tensor *kernel,
                          struct csinn_tensor *bias, struct csinn_conv2d_params *params, int act);
static int conv_init_common(struct csinn_tensor *input, struct csinn_tensor *output,
                            struct csinn_tensor *kernel, struct csinn_tensor *bias,
                            struct csinn_conv2d_params *params, int act)
{
    /* CSINN_QUANT_FLOAT16_W_INT8: fp16 activations over int8 weights */
    if (input->dtype == CSINN_DTYPE_FLOAT16 && kernel->dtype == CSINN_DTYPE_INT8) {
        /* the tcgen05 GEMM path multiplies exact integer weights and scales per channel in the epilogue; depthwise
         * and the shapes that may take the first-layer kernel (K <= 160, O <= 64) keep weights dequantised to fp16 */
        const int grp = params->group > 0 ? params->group : 1;
        const int dwise = grp == input->dim[1] && kernel->dim[1] == 1 && grp > 1;
        const int kd = kernel->dim[1] * kernel->dim[2] * kernel->dim[3];
        const int gemm_only = !dwise && !(grp == 1 && kd <= 160 && kernel->dim[0] <= 64);
        struct csinn_tensor *kf = gemm_only ? b200_int8_weights_as_f16(kernel) : b200_dequant_weights_f16(kernel);
        if (!kf) return CSINN_FALSE;
        const int rc = conv_init_impl(input, output, kf, bias, params, act);
        b200_free_dequant(kf);
        return rc;
    }
    return conv_init_impl(input, output, kernel, bias, params, act);
}
static int conv_init_impl(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_tensor *kernel,
                          struct csinn_tensor *bias, struct csinn_conv2d_params *params, int act)
{
    (void)conv_act_of;
    if (input->dim_count != 4 || kernel->dim_count != 4) {
        b200_fail("conv2d: expected 4-D NCHW input and OIHW kernel");
        return CSINN_UNSUPPORT_LAYOUT;
    }
    if (params->base.layout != CSINN_LAYOUT_NCHW && params->base.layout != 0) {
        b200_fail("conv2d: only CSINN_LAYOUT_NCHW tensors are supported (got %d)", params->base.layout);
        return CSINN_UNSUPPORT_LAYOUT;
    }
    const int C = input->dim[1], O = kernel->dim[0], cg = kernel->dim[1];
    const int kh = kernel->dim[2], kw = kernel->dim[3];
    int group = params->group > 0 ? params->group : 1;
    /* a depthwise convolution with a depth multiplier (O = m * C, source/reference/convolution.c:229: output channel
     * ic * m + j reads input channel ic) is the group convolution group = C, one input channel per group */
    const int is_dw = group == C && cg == 1 && group > 1 && O == C;
    if (!is_dw && (C % group || O % group || cg != C / group)) {
        b200_fail("conv2d: inconsistent group=%d for C=%d O=%d kernel I=%d", group, C, O, cg);
        return CSINN_FALSE;
    }
    int dh = params->dilation_height, dw = params->dilation_width;
    if (kh == 1) dh = 1;
    if (kw == 1) dw = 1;
    if (dh < 1 || dw < 1 || params->stride_height < 1 || params->stride_width < 1) {
        b200_fail("conv2d: stride %dx%d / dilation %dx%d must be >= 1", params->stride_height,
                  params->stride_width, params->dilation_height, params->dilation_width);
        return CSINN_FALSE;
    }
    b200_op *op = op_new(&params->base, is_dw ? B200_OPK_DW : B200_OPK_CONV, input->dtype,
                         is_dw ? "b200_dwconv2d" : "b200_im2col_gemm_tcgen05");
    if (!op) return CSINN_FALSE;
    op->cin = C, op->o = O, op->kh = kh, op->kw = kw;
    op->sh = params->stride_height, op->sw = params->stride_width;
    op->pt = params->pad_top, op->pl = params->pad_left, op->dh = dh, op->dw = dw;
    op->group = is_dw ? 1 : group;
    op->act = act;
    int rc;
    size_t wbytes = 0;
    if (is_dw) {
        const int cp = b200_round_channels(C, op->eb);
        rc = b200_make_requant(op, input, kernel, bias, output, kh * kw, params->conv_extra.fuse_zp2bias, O);
        if (rc == CSINN_TRUE && !(op->d_w = b200_pack_dw_weights(op, kernel, cp, &wbytes))) rc = CSINN_FALSE;
        if (rc == CSINN_TRUE && op->dtype == B200_I8 && kh == 3 && kw == 3 &&
            !(op->d_w2 = b200_pack_dw3x3_rows(op, kernel, cp)))
            rc = CSINN_FALSE;
        /* int8, channels a multiple of 64, symmetric weights: the depthwise convolution can run as an implicit GEMM on
         * the tensor cores against tap-diagonal weights (b200_conv_igemm, dw_slab) */
        if (rc == CSINN_TRUE && op->dtype == B200_I8 && C % 64 == 0 && !op->d_wzp && kh * kw <= 25 && output->dim_count == 4) {
            rc = b200_make_igemm_tables(op, kernel, input->dim[2], input->dim[3], output->dim[2], output->dim[3]);
            if (rc == CSINN_TRUE && op->ig_ncls >= 1 && !(op->d_w_diag = b200_pack_dw_diag(op, kernel, cp, &op->ldk_diag)))
                rc = CSINN_FALSE;
        }
    } else {
        op->kdim = cg * kh * kw;
        op->ldk = b200_round_channels(op->kdim, op->eb);
        op->direct = kh == 1 && kw == 1 && op->sh == 1 && op->sw == 1 && op->pt == 0 && op->pl == 0 &&
                     params->pad_down == 0 && params->pad_right == 0 && group == 1;
        if (op->direct) op->kname = "b200_gemm_tcgen05";
        rc = b200_make_requant(op, input, kernel, bias, output, op->kdim, params->conv_extra.fuse_zp2bias, O);
        if (rc == CSINN_TRUE && !(op->d_w = b200_pack_conv_weights(op, kernel, &wbytes))) rc = CSINN_FALSE;
        /* every convolution that is not a plain GEMM (k > 1, strided or padded 1x1), int8, group 1, channels a
         * multiple of 64, symmetric weights: the implicit-GEMM kernel (K = taps x channels exactly) */
        if (rc == CSINN_TRUE && !op->direct && op->dtype == B200_I8 && group == 1 && C % 64 == 0 && !op->d_wzp &&
            output->dim_count == 4)
            rc = b200_make_igemm_tables(op, kernel, input->dim[2], input->dim[3], output->dim[2], output->dim[3]);
    }
    if (rc != CSINN_TRUE) {
        free(op);
        return rc;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = is_dw ? (int (*)())shl_b200_depthwise_conv2d : (int (*)())shl_b200_conv2d;
    return CSINN_TRUE;
}

int shl_b200_conv2d_init(struct csinn_tensor *input, struct csinn_tensor *output,
                         struct csinn_tensor *kernel, struct csinn_tensor *bias,
                         struct csinn_conv2d_params *params)
{
    return conv_init_common(input, output, kernel, bias, params, B200_ACT_NONE);
}
static int conv2d_relu_init(struct csinn_tensor *input, struct csinn_tensor *output,
                            struct csinn_tensor *kernel, struct csinn_tensor *bias,
                            struct csinn_conv2d_params *params)
{
    return conv_init_common(input, output, kernel, bias, params, B200_ACT_RELU);
}
static int conv2d_relu6_init(struct csinn_tensor *input, struct csinn_tensor *output,
                             struct csinn_tensor *kernel, struct csinn_tensor *bias,
                             struct csinn_conv2d_params *params)
{
    return conv_init_common(input, output, kernel, bias, params, B200_ACT_RELU6);
}
void *shl_b200_conv2d_relu_init_fn(void) { return (void *)conv2d_relu_init; }
void *shl_b200_conv2d_relu6_init_fn(void) { return (void *)conv2d_relu6_init; }

int shl_b200_conv2d(struct csinn_tensor *input, struct csinn_tensor *output,
                    struct csinn_tensor *kernel, struct csinn_tensor *bias,
                    struct csinn_conv2d_params *params)
{
    (void)kernel;
    (void)bias;
    return layer_exec(params, input, NULL, output);
}

int shl_b200_depthwise_conv2d_init(struct csinn_tensor *input, struct csinn_tensor *output,
                                   struct csinn_tensor *kernel, struct csinn_tensor *bias,
                                   struct csinn_conv2d_params *params)
{
    return conv_init_common(input, output, kernel, bias, params, B200_ACT_NONE);
}
int shl_b200_depthwise_conv2d(struct csinn_tensor *input, struct csinn_tensor *output,
                              struct csinn_tensor *kernel, struct csinn_tensor *bias,
                              struct csinn_conv2d_params *params)
{
    (void)kernel;
    (void)bias;
    return layer_exec(params, input, NULL, output);
}

/* ---- fullyconnected ---------------------------------------------------------------------------- */
static int fc_init_impl(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_tensor *weights,
                        struct csinn_tensor *bias, struct csinn_fc_params *params);
int shl_b200_fullyconnected_init(struct csinn_tensor *input, struct csinn_tensor *output,
                                 struct csinn_tensor *weights, struct csinn_tensor *bias,
                                 struct csinn_fc_params *params)
{
    if (input->dtype == CSINN_DTYPE_FLOAT16 && weights->dtype == CSINN_DTYPE_INT8) { /* CSINN_QUANT_FLOAT16_W_INT8 */
        struct csinn_tensor *wf = b200_int8_weights_as_f16(weights); /* exact integers, scale in the GEMM epilogue */
        if (!wf) return CSINN_FALSE;
        const int rc = fc_init_impl(input, output, wf, bias, params);
        b200_free_dequant(wf);
        return rc;
    }
    return fc_init_impl(input, output, weights, bias, params);
}
static int fc_init_impl(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_tensor *weights,
                        struct csinn_tensor *bias, struct csinn_fc_params *params)
{
    if (weights->dim_count != 2) {
        b200_fail("fullyconnected: expected [out][in] weights");
        return CSINN_FALSE;
    }
    b200_dt din;
    if (!b200_dt_from_tensor(&din, input) || din.h * din.w != 1 || din.c != weights->dim[1]) {
        /* source/reference/fullyconnected.c:21 flattens any leading dims in NCHW order; on the
         * pixel-major device layout that only coincides when H*W == 1 */
        b200_fail("fullyconnected: input must be [batch][%d] (or N x %d x 1 x 1)", weights->dim[1], weights->dim[1]);
        return CSINN_FALSE;
    }
    b200_op *op = op_new(&params->base, B200_OPK_FC, input->dtype, "b200_gemm_tcgen05");
    if (!op) return CSINN_FALSE;
    op->o = weights->dim[0], op->cin = weights->dim[1], op->kdim = weights->dim[1];
    op->ldk = b200_round_channels(op->kdim, op->eb);
    op->group = 1, op->direct = 1;
    size_t wbytes = 0;
    int rc = b200_make_requant(op, input, weights, bias, output, op->kdim, params->fc_extra.fuse_zp2bias, op->o);
    if (rc == CSINN_TRUE && !(op->d_w = b200_pack_fc_weights(op, weights, &wbytes))) rc = CSINN_FALSE;
    if (rc != CSINN_TRUE) {
        free(op);
        return rc;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())shl_b200_fullyconnected;
    return CSINN_TRUE;
}
int shl_b200_fullyconnected(struct csinn_tensor *input, struct csinn_tensor *output,
                            struct csinn_tensor *weights, struct csinn_tensor *bias,
                            struct csinn_fc_params *params)
{
    (void)weights;
    (void)bias;
    return layer_exec(params, input, NULL, output);
}

/* ---- relu / relu6 -------------------------------------------------------------------------------- */
static const char *const kActNames[] = {"b200_identity", "b200_relu", "b200_relu6", "b200_leaky_relu", "b200_sigmoid",
                                        "b200_clip", "b200_silu", "b200_erf"};

static int act_init_p(struct csinn_tensor *input, struct csinn_tensor *output, void *params, int act, float p0, float p1)
{
    struct csinn_params_base *base = params;
    b200_op *op = op_new(base, B200_OPK_ACT, input->dtype, kActNames[act]);
    if (!op) return CSINN_FALSE;
    op->act = act, op->act_p0 = p0, op->act_p1 = p1;
    if (op->dtype == B200_I8) {
        if (!input->qinfo || !output->qinfo) {
            b200_fail("relu: int8 tensors without qinfo");
            free(op);
            return CSINN_FALSE;
        }
        op->d_lut = upload_lut(op->ctx, act, p0, p1, input->qinfo->scale, input->qinfo->zero_point,
                               output->qinfo->scale, output->qinfo->zero_point);
        if (!op->d_lut) {
            free(op);
            return CSINN_FALSE;
        }
    }
    b200_op_bind(params, op);
    base->cb->exec = (int (*)())shl_b200_relu;
    return CSINN_TRUE;
}
static int act_init(struct csinn_tensor *input, struct csinn_tensor *output, void *params, int act)
{
    return act_init_p(input, output, params, act, 0.f, 0.f);
}
/* leaky relu / sigmoid / clip: replace shl_rvv_leaky_relu_int8, shl_rvv_sigmoid_*, shl_rvv_clip_int8
 * (source/thead_rvv/setup.c:154-508 registrations); semantics source/reference/leaky_relu.c:33,
 * sigmoid.c:33, clip.c:32-38 through shl_ref_siso_callback_base */
static int leaky_relu_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_relu_params *params)
{
    return act_init_p(input, output, params, B200_ACT_LEAKY_RELU, params->n, 0.f);
}
static int sigmoid_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_sigmoid_params *params)
{
    return act_init_p(input, output, params, B200_ACT_SIGMOID, 0.f, 0.f);
}
static int clip_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_clip_params *params)
{
    return act_init_p(input, output, params, B200_ACT_CLIP, params->min_value, params->max_value);
}
/* silu / erf: replace the shl_gref_silu / shl_gref_erf registrations of source/thead_rvv/setup.c (which run
 * source/reference/silu.c:37, erf.c:37) */
static int silu_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_sigmoid_params *params)
{
    return act_init_p(input, output, params, B200_ACT_SILU, 0.f, 0.f);
}
static int erf_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_siso_params *params)
{
    return act_init_p(input, output, params, B200_ACT_ERF, 0.f, 0.f);
}
void *shl_b200_silu_init_fn(void) { return (void *)silu_init; }
void *shl_b200_erf_init_fn(void) { return (void *)erf_init; }
void *shl_b200_leaky_relu_init_fn(void) { return (void *)leaky_relu_init; }
void *shl_b200_sigmoid_init_fn(void) { return (void *)sigmoid_init; }
void *shl_b200_clip_init_fn(void) { return (void *)clip_init; }
int shl_b200_relu_init(struct csinn_tensor *input, struct csinn_tensor *output,
                       struct csinn_relu_params *params)
{
    return act_init(input, output, params, B200_ACT_RELU);
}
static int relu6_init(struct csinn_tensor *input, struct csinn_tensor *output,
                      struct csinn_relu_params *params)
{
    return act_init(input, output, params, B200_ACT_RELU6);
}
void *shl_b200_relu6_init_fn(void) { return (void *)relu6_init; }
int shl_b200_relu(struct csinn_tensor *input, struct csinn_tensor *output,
                  struct csinn_relu_params *params)
{
    return layer_exec(params, input, NULL, output);
}

/* ---- add -------------------------------------------------------------------------------------------- */
static int binary_init(struct csinn_tensor *input0, struct csinn_tensor *input1, struct csinn_tensor *output,
                       struct csinn_diso_params *params, int binop);
int shl_b200_add_init(struct csinn_tensor *input0, struct csinn_tensor *input1,
                      struct csinn_tensor *output, struct csinn_diso_params *params)
{
    return binary_init(input0, input1, output, params, B200_BINOP_ADD);
}
/* sub / mul: replace shl_rvv_sub_int8 / shl_rvv_mul_int8 (source/thead_rvv/setup.c registrations) for
 * same-shape operands; exec is shl_b200_add */
static int sub_init(struct csinn_tensor *input0, struct csinn_tensor *input1, struct csinn_tensor *output,
                    struct csinn_diso_params *params)
{
    return binary_init(input0, input1, output, params, B200_BINOP_SUB);
}
static int mul_init(struct csinn_tensor *input0, struct csinn_tensor *input1, struct csinn_tensor *output,
                    struct csinn_diso_params *params)
{
    return binary_init(input0, input1, output, params, B200_BINOP_MUL);
}
/* prelu: replaces the shl_gref_prelu registration of source/thead_rvv/setup.c (source/reference/prelu.c:56):
 * a binary op between the input and the constant per-channel slope, slope indexed along the channel axis */
static int prelu_init(struct csinn_tensor *input, struct csinn_tensor *alpha, struct csinn_tensor *output,
                      struct csinn_prelu_params *params)
{
    const int ch_axis = input->dim_count >= 2 ? 1 : 0;
    if (!alpha->is_const || params->axis != ch_axis) {
        b200_fail("prelu: the slope must be a constant indexed along the channel axis (axis %d given)", params->axis);
        return CSINN_FALSE;
    }
    return binary_init(input, alpha, output, (struct csinn_diso_params *)params, B200_BINOP_PRELU);
}
void *shl_b200_prelu_init_fn(void) { return (void *)prelu_init; }
/* div: replaces the shl_gref_div registration of source/thead_rvv/setup.c (source/reference/div.c:36) */
static int div_init(struct csinn_tensor *input0, struct csinn_tensor *input1, struct csinn_tensor *output,
                    struct csinn_diso_params *params)
{
    return binary_init(input0, input1, output, params, B200_BINOP_DIV);
}
void *shl_b200_div_init_fn(void) { return (void *)div_init; }
void *shl_b200_sub_init_fn(void) { return (void *)sub_init; }
void *shl_b200_mul_init_fn(void) { return (void *)mul_init; }
static int binary_init(struct csinn_tensor *input0, struct csinn_tensor *input1, struct csinn_tensor *output,
                       struct csinn_diso_params *params, int binop)
{
    /* a constant second operand may be one element or one value per channel ([C], [C,1,1], [1,C,1,1]):
     * the broadcasting of shl_ref_add_f32 (source/reference/add.c:21); anything else must match in0 */
    int per_channel = 0, scalar = 0, bcast_nc = 0;
    b200_dt d0;
    if (input0->is_const || !b200_dt_from_tensor(&d0, input0)) {
        b200_fail("add: first operand must be a non-constant int8 / fp16 tensor of rank 1..4");
        return CSINN_FALSE;
    }
    if (input1->is_const) {
        int64_t elems = 1;
        int big = 0, big_dim = -1;
        for (int i = 0; i < input1->dim_count; i++) {
            elems *= input1->dim[i];
            if (input1->dim[i] != 1) big++, big_dim = i;
        }
        const int from_end = input1->dim_count - 1 - big_dim; /* [.., C, 1, 1] against NCHW: the channel axis */
        const int ch_from_end = input0->dim_count - 1 - (input0->dim_count >= 2 ? 1 : 0);
        scalar = elems == 1;
        per_channel = !scalar && big == 1 && elems == d0.c && from_end == ch_from_end;
        /* prelu: the slope tensor is [C] whatever the rank of the input (indexed along params->axis) */
        if (binop == B200_BINOP_PRELU) per_channel = !scalar && elems == d0.c, scalar = scalar && d0.c == 1;
        if (!input1->data || input1->dtype != input0->dtype || (!scalar && !per_channel)) {
            b200_fail("add: constant second operand must be one element or one value per channel of the same dtype");
            return CSINN_FALSE;
        }
    } else {
        /* two activations: the same shape, or [N, C, H, W] against [N, C, 1, 1] (one value per image and
         * channel -- a squeeze-and-excitation scale; source/reference/utils.c:83) */
        int same = input0->dim_count == input1->dim_count;
        for (int i = 0; same && i < input0->dim_count; i++) same = input0->dim[i] == input1->dim[i];
        if (!same) {
            const int ok = input0->dim_count == 4 && input1->dim_count == 4 && input1->dim[1] == input0->dim[1] &&
                           input1->dim[2] == 1 && input1->dim[3] == 1 && input1->dim[0] == input0->dim[0] &&
                           input1->dtype == input0->dtype && binop != B200_BINOP_PRELU;
            if (!ok) {
                b200_fail("add: broadcasting between two activations is supported for [N, C, H, W] against [N, C, 1, 1] only");
                return CSINN_FALSE;
            }
            bcast_nc = 1;
        }
    }
    static const char *const names[] = {"b200_add", "b200_sub", "b200_mul", "b200_prelu", "b200_div"};
    b200_op *op = op_new(&params->base, B200_OPK_ADD, input0->dtype, names[binop]);
    if (!op) return CSINN_FALSE;
    op->binop = binop;
    op->bcast_nc = bcast_nc;
    if (input1->is_const) {
        /* one pixel's worth of channels, padding lanes 0 (their results are never read) */
        uint8_t *row = calloc((size_t)d0.cp, (size_t)op->eb);
        if (!row) {
            free(op);
            return CSINN_FALSE;
        }
        for (int c = 0; c < d0.c; c++)
            memcpy(row + (size_t)c * op->eb, (const uint8_t *)input1->data + (size_t)(scalar ? 0 : c) * op->eb, (size_t)op->eb);
        op->d_const = b200_warena_put(op->ctx, row, (size_t)d0.cp * op->eb);
        op->const_count = d0.cp;
        free(row);
        if (!op->d_const) {
            free(op);
            return CSINN_FALSE;
        }
    }
    if (op->dtype == B200_I8) {
        if (!input0->qinfo || !input1->qinfo || !output->qinfo) {
            b200_fail("add: int8 tensors without qinfo");
            free(op);
            return CSINN_FALSE;
        }
        op->s_in = input0->qinfo->scale, op->zp_in = input0->qinfo->zero_point;
        op->s_in1 = input1->qinfo->scale, op->zp_in1 = input1->qinfo->zero_point;
        op->s_out = output->qinfo->scale, op->zp_out = output->qinfo->zero_point;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())shl_b200_add;
    return CSINN_TRUE;
}
int shl_b200_add(struct csinn_tensor *input0, struct csinn_tensor *input1,
                 struct csinn_tensor *output, struct csinn_diso_params *params)
{
    return layer_exec(params, input0, input1->is_const ? NULL : input1, output);
}

/* ---- pooling ------------------------------------------------------------------------------------------ */
static int pool_init_common(struct csinn_tensor *input, struct csinn_tensor *output,
                            struct csinn_pool_params *params, int avg, int global)
{
    if (input->dim_count != 4) {
        b200_fail("pool2d: expected a 4-D NCHW input");
        return CSINN_UNSUPPORT_LAYOUT;
    }
    b200_op *op = op_new(&params->base, B200_OPK_POOL, input->dtype,
                         global ? (avg ? "b200_global_avgpool" : "b200_global_maxpool") : (avg ? "b200_avgpool" : "b200_maxpool"));
    if (!op) return CSINN_FALSE;
    op->pool_avg = avg, op->pool_global = global;
    op->kh = params->filter_height, op->kw = params->filter_width;
    op->sh = params->stride_height, op->sw = params->stride_width;
    op->pt = params->pad_top, op->pl = params->pad_left;
    op->count_include_pad = params->count_include_pad ? 1 : 0;
    if (!global && (op->kh < 1 || op->kw < 1 || op->sh < 1 || op->sw < 1)) {
        b200_fail("pool2d: filter %dx%d / stride %dx%d must be >= 1", op->kh, op->kw, op->sh, op->sw);
        free(op);
        return CSINN_FALSE;
    }
    if (op->dtype == B200_I8) {
        if (!input->qinfo || !output->qinfo) {
            b200_fail("pool2d: int8 tensors without qinfo");
            free(op);
            return CSINN_FALSE;
        }
        op->s_in = input->qinfo->scale, op->zp_in = input->qinfo->zero_point;
        op->s_out = output->qinfo->scale, op->zp_out = output->qinfo->zero_point;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())shl_b200_pool2d;
    return CSINN_TRUE;
}
int shl_b200_pool2d_init(struct csinn_tensor *input, struct csinn_tensor *output,
                         struct csinn_pool_params *params)
{
    return pool_init_common(input, output, params, 0, 0); /* maxpool */
}
static int avgpool_init(struct csinn_tensor *input, struct csinn_tensor *output,
                        struct csinn_pool_params *params)
{
    return pool_init_common(input, output, params, 1, 0);
}
static int global_avgpool_init(struct csinn_tensor *input, struct csinn_tensor *output,
                               struct csinn_pool_params *params)
{
    return pool_init_common(input, output, params, 1, 1);
}
/* replaces shl_rvv_global_maxpool2d_init_int8 / _fp16 (source/thead_rvv/setup.c); semantics
 * source/reference/global_maxpool.c:21: max pooling over the whole map, stride 1, no padding */
static int global_maxpool_init(struct csinn_tensor *input, struct csinn_tensor *output,
                               struct csinn_pool_params *params)
{
    return pool_init_common(input, output, params, 0, 1);
}
void *shl_b200_global_maxpool_init_fn(void) { return (void *)global_maxpool_init; }
void *shl_b200_avgpool_init_fn(void) { return (void *)avgpool_init; }
void *shl_b200_global_avgpool_init_fn(void) { return (void *)global_avgpool_init; }
int shl_b200_pool2d(struct csinn_tensor *input, struct csinn_tensor *output,
                    struct csinn_pool_params *params)
{
    return layer_exec(params, input, NULL, output);
}

/* ---- softmax -------------------------------------------------------------------------------------------- */
int shl_b200_softmax_init(struct csinn_tensor *input, struct csinn_tensor *output,
                          struct csinn_softmax_params *params)
{
    b200_dt din;
    int axis = params->axis < 0 ? params->axis + input->dim_count : params->axis;
    /* axis 1 = the channel axis: on the pixel-major device layout every (image, y, x) position's channels are one contiguous
     * row, so an N x C x H x W softmax is N * H * W rows of the [N][C] kernel (the reference walks them in the same order,
     * source/reference/softmax.c:42-63) */
    if (!b200_dt_from_tensor(&din, input) || axis != 1) {
        b200_fail("softmax: only axis 1 (channels) of a rank 2..4 tensor is supported");
        return CSINN_FALSE;
    }
    b200_op *op = op_new(&params->base, B200_OPK_SOFTMAX, input->dtype, "b200_softmax");
    if (!op) return CSINN_FALSE;
    if (op->dtype == B200_I8) {
        if (!input->qinfo || !output->qinfo) {
            b200_fail("softmax: int8 tensors without qinfo");
            free(op);
            return CSINN_FALSE;
        }
        op->s_in = input->qinfo->scale, op->zp_in = input->qinfo->zero_point;
        op->s_out = output->qinfo->scale, op->zp_out = output->qinfo->zero_point;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())shl_b200_softmax;
    return CSINN_TRUE;
}
int shl_b200_softmax(struct csinn_tensor *input, struct csinn_tensor *output,
                     struct csinn_softmax_params *params)
{
    return layer_exec(params, input, NULL, output);
}

/* ---- reshape / flatten ------------------------------------------------------------------------------------ */
int shl_b200_reshape_init(struct csinn_tensor *input, struct csinn_tensor *output, void *params)
{
    struct csinn_params_base *base = params;
    b200_dt din, dout;
    if (!b200_dt_from_tensor(&din, input) || !b200_dt_from_tensor(&dout, output)) {
        b200_fail("reshape/flatten: unsupported tensor rank / dtype");
        return CSINN_FALSE;
    }
    /* N x C x 1 x 1 <-> N x C: the pixel-major buffer already is the result.  N x C x H x W -> N x (C*H*W) (the
     * flatten in front of a VGG / AlexNet classifier; source/reference/flatten.c / reshape.c copy the NCHW bytes):
     * the pixel-major buffer is permuted back into NCHW order, which IS the flattened row when C*H*W fills the
     * output's row pitch.  Other reshapes take the same permutation there and a second one back (direct = 2). */
    const int plain = din.h * din.w == 1 && dout.h * dout.w == 1 && din.n == dout.n && din.c == dout.c;
    const long long flat = (long long)din.c * din.h * din.w;
    const int to_rows = !plain && dout.h * dout.w == 1 && din.n == dout.n && flat == dout.c && dout.cp == dout.c;
    const long long e_in = (long long)din.n * din.c * din.h * din.w, e_out = (long long)dout.n * dout.c * dout.h * dout.w;
    if (e_in != e_out) {
        b200_fail("reshape/flatten: %lld elements in, %lld out", e_in, e_out);
        return CSINN_FALSE;
    }
    /* anything else goes through the NCHW order in a scratch buffer (two layout kernels): slow, general */
    const int general = !plain && !to_rows;
    b200_op *op = op_new(base, B200_OPK_COPY, input->dtype,
                         to_rows ? "b200_flatten_to_nchw" : (general ? "b200_reshape_permute" : "b200_reshape_copy"));
    if (!op) return CSINN_FALSE;
    op->direct = to_rows ? 1 : (general ? 2 : 0); /* B200_OPK_COPY: 1 = permute pixel-major -> NCHW order, 2 = there and back */
    b200_op_bind(params, op);
    base->cb->exec = (int (*)())shl_b200_reshape;
    return CSINN_TRUE;
}
int shl_b200_reshape(struct csinn_tensor *input, struct csinn_tensor *output, void *params)
{
    return layer_exec(params, input, NULL, output);
}

/* ---- concat --------------------------------------------------------------------------------------------- */
/* replaces shl_rvv_concat_int8 / _fp16 (source/thead_rvv/setup.c registrations); semantics
 * source/reference/concat.c:52-80: each input dequantised with its own qinfo, the output quantised
 * with its own -> one 256-entry table per input.  Any axis of a rank 1..4 tensor. */
int shl_b200_concat_init(struct csinn_tensor **input, struct csinn_tensor *output,
                         struct csinn_concat_params *params)
{
    const int k = params->inputs_count;
    if (k < 1 || k > B200_CONCAT_MAX) {
        b200_fail("concat: %d inputs (1..%d supported)", k, B200_CONCAT_MAX);
        return CSINN_FALSE;
    }
    int axis = params->axis;
    if (axis < 0) axis += output->dim_count;
    if (axis < 0 || axis >= output->dim_count) {
        b200_fail("concat: axis %d of a rank-%d tensor", params->axis, output->dim_count);
        return CSINN_FALSE;
    }
    /* API axis -> device axis of b200_dt_from_tensor's (n, c, h, w) view */
    static const int dev_axis[5][4] = {{0}, {1}, {0, 1}, {0, 1, 3}, {0, 1, 2, 3}};
    b200_dt dout;
    if (!b200_dt_from_tensor(&dout, output)) {
        b200_fail("concat: unsupported output tensor (dtype %d, rank %d)", output->dtype, output->dim_count);
        return CSINN_FALSE;
    }
    b200_op *op = op_new(&params->base, B200_OPK_CONCAT, output->dtype, "b200_concat_slice");
    if (!op) return CSINN_FALSE;
    op->cat_n = k, op->cat_axis = dev_axis[output->dim_count][axis];
    int along = 0;
    for (int i = 0; i < k; i++) {
        const struct csinn_tensor *t = input[i];
        int ok = t && t->dtype == output->dtype && t->dim_count == output->dim_count && !t->is_const;
        for (int d = 0; ok && d < output->dim_count; d++)
            if (d != axis && t->dim[d] != output->dim[d]) ok = 0;
        if (ok && op->dtype == B200_I8 && (!t->qinfo || !output->qinfo)) ok = 0;
        if (!ok) {
            b200_fail("concat: input %d does not match the output (dtype, rank, the other dimensions, qinfo) or is "
                      "a constant", i);
            free(op);
            return CSINN_FALSE;
        }
        op->cat_off[i] = along;
        along += t->dim[axis];
        if (op->dtype == B200_I8) {
            int8_t lut[256];
            b200_build_requant_lut(lut, B200_ACT_NONE, t->qinfo->scale, t->qinfo->zero_point, output->qinfo->scale,
                                   output->qinfo->zero_point);
            op->cat_lut[i] = b200_warena_put(op->ctx, lut, 256);
            if (!op->cat_lut[i]) {
                free(op);
                return CSINN_FALSE;
            }
        }
    }
    if (along != output->dim[axis]) {
        b200_fail("concat: inputs add up to %d along axis %d, the output has %d", along, axis, output->dim[axis]);
        free(op);
        return CSINN_FALSE;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())shl_b200_concat;
    return CSINN_TRUE;
}

int shl_b200_concat(struct csinn_tensor **input, struct csinn_tensor *output, struct csinn_concat_params *params)
{
    b200_op *op = b200_op_find(params);
    if (!op || op->kind != B200_OPK_CONCAT) {
        b200_fail("concat exec before init: no b200 operator bound to these params");
        return CSINN_FALSE;
    }
    b200_set_device(op->ctx->device);
    void *stream = op->ctx->stream;
    b200_dt dout;
    if (!b200_dt_from_tensor(&dout, output) || !output->data) {
        b200_fail("unsupported or unallocated output tensor");
        return CSINN_FALSE;
    }
    const size_t raw = (size_t)dout.n * dout.c * dout.h * dout.w * dout.eb;
    dout.d = stage(op, 4, b200_dt_bytes(&dout));
    void *d_raw = stage(op, 5, raw);
    if (!dout.d || !d_raw) return CSINN_FALSE;
    for (int i = 0; i < op->cat_n; i++) {
        /* the staging slots are reused input after input: everything is ordered on one stream */
        b200_dt din;
        if (upload_nchw(op, 0, input[i], &din, stream) != CSINN_TRUE) return CSINN_FALSE;
        if (b200_op_run(op, i, &din, NULL, &dout, NULL, stream) != CSINN_TRUE) return CSINN_FALSE;
    }
    DEV_CHECK(b200_nhwc_to_nchw(dout.d, d_raw, dout.n, dout.c, dout.h, dout.w, dout.cp, dout.eb, stream));
    DEV_CHECK(b200_memcpy_d2h(output->data, d_raw, raw, stream));
    DEV_CHECK(b200_stream_sync(stream));
    return CSINN_TRUE;
}

/* ---- split ---------------------------------------------------------------------------------------------- */
/* replaces the shl_gref_split registration of source/thead_rvv/setup.c (which runs source/reference/split.c:81):
 * output i is the slice [split_index[i-1], split_index[i]) of the input along `axis` (equal chunks when
 * split_index is NULL, the last one shorter), requantised with its own qinfo -> one table per output */
int shl_b200_split_init(struct csinn_tensor *input, struct csinn_tensor **output, struct csinn_split_params *params)
{
    const int k = params->output_num;
    if (k < 1 || k > B200_CONCAT_MAX) {
        b200_fail("split: %d outputs (1..%d supported)", k, B200_CONCAT_MAX);
        return CSINN_FALSE;
    }
    int axis = params->axis;
    if (axis < 0) axis += input->dim_count;
    b200_dt din;
    if (axis < 0 || axis >= input->dim_count || !b200_dt_from_tensor(&din, input)) {
        b200_fail("split: axis %d of a rank-%d tensor (dtype %d)", params->axis, input->dim_count, input->dtype);
        return CSINN_FALSE;
    }
    static const int dev_axis[5][4] = {{0}, {1}, {0, 1}, {0, 1, 3}, {0, 1, 2, 3}};
    b200_op *op = op_new(&params->base, B200_OPK_SPLIT, input->dtype, "b200_split_slice");
    if (!op) return CSINN_FALSE;
    op->cat_n = k, op->cat_axis = dev_axis[input->dim_count][axis];
    const int avg = (input->dim[axis] + k - 1) / k;
    for (int i = 0; i < k; i++) {
        const int begin = params->split_index ? (i == 0 ? 0 : params->split_index[i - 1]) : i * avg;
        const int end = (i == k - 1) ? input->dim[axis] : (params->split_index ? params->split_index[i] : (i + 1) * avg);
        const struct csinn_tensor *t = output[i];
        int ok = t && t->dtype == input->dtype && t->dim_count == input->dim_count && begin >= 0 && end > begin &&
                 end <= input->dim[axis] && t->dim[axis] == end - begin;
        for (int d = 0; ok && d < input->dim_count; d++)
            if (d != axis && t->dim[d] != input->dim[d]) ok = 0;
        if (ok && op->dtype == B200_I8 && (!t->qinfo || !input->qinfo)) ok = 0;
        if (!ok) {
            b200_fail("split: output %d does not match slice [%d, %d) of the input (dtype, rank, dimensions, qinfo)", i,
                      begin, end);
            free(op);
            return CSINN_FALSE;
        }
        op->cat_off[i] = begin;
        if (op->dtype == B200_I8) {
            int8_t lut[256];
            b200_build_requant_lut(lut, B200_ACT_NONE, input->qinfo->scale, input->qinfo->zero_point, t->qinfo->scale,
                                   t->qinfo->zero_point);
            op->cat_lut[i] = b200_warena_put(op->ctx, lut, 256);
            if (!op->cat_lut[i]) {
                free(op);
                return CSINN_FALSE;
            }
        }
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())shl_b200_split;
    return CSINN_TRUE;
}

int shl_b200_split(struct csinn_tensor *input, struct csinn_tensor **output, struct csinn_split_params *params)
{
    b200_op *op = b200_op_find(params);
    if (!op || op->kind != B200_OPK_SPLIT) {
        b200_fail("split exec before init: no b200 operator bound to these params");
        return CSINN_FALSE;
    }
    b200_set_device(op->ctx->device);
    void *stream = op->ctx->stream;
    b200_dt din;
    if (upload_nchw(op, 0, input, &din, stream) != CSINN_TRUE) return CSINN_FALSE;
    for (int i = 0; i < op->cat_n; i++) {
        b200_dt dout;
        if (!b200_dt_from_tensor(&dout, output[i]) || !output[i]->data) {
            b200_fail("split: unsupported or unallocated output %d", i);
            return CSINN_FALSE;
        }
        const size_t raw = (size_t)dout.n * dout.c * dout.h * dout.w * dout.eb;
        dout.d = stage(op, 4, b200_dt_bytes(&dout));
        void *d_raw = stage(op, 5, raw);
        if (!dout.d || !d_raw) return CSINN_FALSE;
        if (b200_op_run(op, i, &din, NULL, &dout, NULL, stream) != CSINN_TRUE) return CSINN_FALSE;
        DEV_CHECK(b200_nhwc_to_nchw(dout.d, d_raw, dout.n, dout.c, dout.h, dout.w, dout.cp, dout.eb, stream));
        DEV_CHECK(b200_memcpy_d2h(output[i]->data, d_raw, raw, stream));
        DEV_CHECK(b200_stream_sync(stream)); /* the staging slots are reused by the next output */
    }
    return CSINN_TRUE;
}

/* ---- perf callbacks: kernel name for the trace profiler ------------------------------------------------ */
/* gref calls perf with the op's own argument list plus a trailing csinn_perf_info*
 * (source/graph_ref/setup.c:509-540), hence one function per arity. */
static int perf_set(void *params, struct csinn_perf_info *perf_info)
{
    b200_op *op = b200_op_find(params);
    if (perf_info) perf_info->kernel_name = (char *)(op ? op->kname : "b200");
    return CSINN_TRUE;
}
int shl_b200_perf(struct csinn_tensor *input, struct csinn_tensor *output,
                  struct csinn_tensor *kernel, struct csinn_tensor *bias, void *params,
                  struct csinn_perf_info *perf_info)
{
    (void)input, (void)output, (void)kernel, (void)bias;
    return perf_set(params, perf_info);
}
int shl_b200_perf_siso(struct csinn_tensor *input, struct csinn_tensor *output, void *params,
                       struct csinn_perf_info *perf_info)
{
    (void)input, (void)output;
    return perf_set(params, perf_info);
}
int shl_b200_perf_diso(struct csinn_tensor *input0, struct csinn_tensor *input1,
                       struct csinn_tensor *output, void *params,
                       struct csinn_perf_info *perf_info)
{
    (void)input0, (void)input1, (void)output;
    return perf_set(params, perf_info);
}

/* ---- transpose / gather / reduce_sum / layer_norm / rms_norm / matmul (source/thead_rvv/setup.c:316-470) ---------- */
static b200_op *tensor_op_new(struct csinn_params_base *base, const struct csinn_tensor *input,
                              const struct csinn_tensor *output, int t_op, const char *kname)
{
    if (input->dim_count < 1 || input->dim_count > 4 || output->dim_count < 1 || output->dim_count > 4) {
        b200_fail("%s: tensors of rank 1..4 only (got %d -> %d)", kname, input->dim_count, output->dim_count);
        return NULL;
    }
    b200_op *op = op_new(base, B200_OPK_TENSOR, input->dtype, kname);
    if (!op) return NULL;
    op->t_op = t_op;
    op->in_rank = input->dim_count, op->out_rank = output->dim_count;
    for (int i = 0; i < input->dim_count; i++) op->in_dim[i] = input->dim[i];
    for (int i = 0; i < output->dim_count; i++) op->out_dim[i] = output->dim[i];
    if (op->dtype == B200_I8) {
        if (!input->qinfo || !output->qinfo) {
            b200_fail("%s: int8 tensors without qinfo", kname);
            free(op);
            return NULL;
        }
        op->s_in = input->qinfo->scale, op->zp_in = input->qinfo->zero_point;
        op->s_out = output->qinfo->scale, op->zp_out = output->qinfo->zero_point;
    }
    return op;
}

/* int8 copies: requant_out(dequant_in(q)) as a table when the two qinfos differ (what the reference's
 * siso_callback_base does around the f32 copy), else a plain copy */
static int tensor_op_copy_lut(b200_op *op)
{
    if (op->dtype != B200_I8 || (op->s_in == op->s_out && op->zp_in == op->zp_out)) return CSINN_TRUE;
    op->d_lut = upload_lut(op->ctx, B200_ACT_NONE, 0.f, 0.f, op->s_in, op->zp_in, op->s_out, op->zp_out);
    return op->d_lut ? CSINN_TRUE : CSINN_FALSE;
}

static int tensor_exec1(struct csinn_tensor *input, struct csinn_tensor *output, void *params)
{
    return layer_exec(params, input, NULL, output);
}

static int transpose_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_transpose_params *params)
{
    b200_op *op = tensor_op_new(&params->base, input, output, B200_T_TRANSPOSE, "b200_transpose");
    if (!op) return CSINN_FALSE;
    if (params->permute_num != input->dim_count || output->dim_count != input->dim_count) {
        b200_fail("transpose: %d permutation entries for a rank-%d tensor", params->permute_num, input->dim_count);
        free(op);
        return CSINN_FALSE;
    }
    for (int k = 0; k < params->permute_num; k++) op->t_perm[k] = params->permute[k];
    if (tensor_op_copy_lut(op) != CSINN_TRUE) {
        free(op);
        return CSINN_FALSE;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())tensor_exec1;
    return CSINN_TRUE;
}

static int gather_exec(struct csinn_tensor *input, struct csinn_tensor *indices, struct csinn_tensor *output, void *params)
{
    (void)indices;
    return layer_exec(params, input, NULL, output);
}
static int gather_init(struct csinn_tensor *input, struct csinn_tensor *indices, struct csinn_tensor *output,
                       struct csinn_gather_params *params)
{
    if (!indices->is_const || !indices->data || (indices->dtype != CSINN_DTYPE_INT64 && indices->dtype != CSINN_DTYPE_INT32)) {
        b200_fail("gather: the indices must be a constant int64 / int32 tensor");
        return CSINN_FALSE;
    }
    b200_op *op = tensor_op_new(&params->base, input, output, B200_T_GATHER, "b200_gather");
    if (!op) return CSINN_FALSE;
    op->t_axis = params->axis < 0 ? params->axis + input->dim_count : params->axis;
    int64_t n = 1;
    for (int i = 0; i < indices->dim_count; i++) n *= indices->dim[i];
    if (op->t_axis < 0 || op->t_axis >= input->dim_count || n <= 0 || n > (1 << 24)) {
        b200_fail("gather: axis %d / %lld indices", params->axis, (long long)n);
        free(op);
        return CSINN_FALSE;
    }
    int32_t *idx = malloc((size_t)n * sizeof(int32_t));
    if (!idx) {
        free(op);
        return CSINN_FALSE;
    }
    for (int64_t i = 0; i < n; i++) {
        const int64_t v = indices->dtype == CSINN_DTYPE_INT64 ? ((const int64_t *)indices->data)[i] : ((const int32_t *)indices->data)[i];
        idx[i] = v > INT32_MAX ? INT32_MAX : (v < INT32_MIN / 2 ? INT32_MIN / 2 : (int32_t)v); /* out of range stays out of range */
    }
    op->d_idx = b200_warena_put(op->ctx, idx, (size_t)n * sizeof(int32_t));
    op->n_idx = (int)n;
    free(idx);
    if (op->dtype == B200_I8) {
        int8_t one[256];
        b200_build_unary_lut(one, B200_ACT_NONE, 0.f, 0.f, 1.0f, 0, op->s_out, op->zp_out); /* quantise 0.0: index of q = 0 */
        op->oob_q = one[128];
    }
    if (!op->d_idx || tensor_op_copy_lut(op) != CSINN_TRUE) {
        free(op);
        return CSINN_FALSE;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())gather_exec;
    return CSINN_TRUE;
}

static int reduce_sum_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_reduce_params *params)
{
    if (params->axis_count != 1) { /* the reference asserts the same, source/reference/reduce_sum.c:25 */
        b200_fail("reduce_sum: exactly one axis (got %d)", params->axis_count);
        return CSINN_FALSE;
    }
    b200_op *op = tensor_op_new(&params->base, input, output, B200_T_REDUCE_SUM, "b200_reduce_sum");
    if (!op) return CSINN_FALSE;
    op->t_axis = params->axis[0]; /* -1 = over everything, as in the reference */
    if (op->t_axis < -1 || op->t_axis >= input->dim_count) {
        b200_fail("reduce_sum: axis %d of a rank-%d tensor", op->t_axis, input->dim_count);
        free(op);
        return CSINN_FALSE;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())tensor_exec1;
    return CSINN_TRUE;
}

/* gamma / beta / weight constants as dequantised f32 (what shl_ref_tensor_transform_f32 hands the f32 kernels) */
static float *const_as_f32(b200_op *op, const struct csinn_tensor *t, int64_t want, const char *what)
{
    int64_t n = 1;
    for (int i = 0; i < t->dim_count; i++) n *= t->dim[i];
    if (!t->data || n != want) {
        b200_fail("%s: constant of %lld elements expected (got %lld)", what, (long long)want, (long long)n);
        return NULL;
    }
    float *f = malloc((size_t)n * sizeof(float));
    if (!f) return NULL;
    for (int64_t i = 0; i < n; i++) {
        if (t->dtype == CSINN_DTYPE_INT8)
            f[i] = ((float)((const int8_t *)t->data)[i] - (float)t->qinfo->zero_point) * t->qinfo->scale;
        else if (t->dtype == CSINN_DTYPE_FLOAT16)
            f[i] = b200_f16_to_f32(((const uint16_t *)t->data)[i]) *
                   (t->qinfo && fabsf(t->qinfo->scale - 1.f) > 1.1920929e-7f ? t->qinfo->scale : 1.f);
        else if (t->dtype == CSINN_DTYPE_FLOAT32)
            f[i] = ((const float *)t->data)[i];
        else {
            b200_fail("%s: constant dtype %d", what, t->dtype);
            free(f);
            return NULL;
        }
    }
    float *d = b200_warena_put(op->ctx, f, (size_t)n * sizeof(float));
    free(f);
    return d;
}

static int norm_exec4(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_tensor *g, struct csinn_tensor *b,
                      void *params)
{
    (void)g, (void)b;
    return layer_exec(params, input, NULL, output);
}
static int layer_norm_init(struct csinn_tensor *input, struct csinn_tensor *output, struct csinn_tensor *gamma,
                           struct csinn_tensor *beta, struct csinn_layer_norm_params *params)
{
    b200_op *op = tensor_op_new(&params->base, input, output, B200_T_LAYER_NORM, "b200_layer_norm");
    if (!op) return CSINN_FALSE;
    op->t_axis = params->axis >= 0 ? params->axis : params->axis + input->dim_count;
    op->t_eps = params->epsilon;
    int64_t norm = 1;
    for (int i = op->t_axis; i >= 0 && i < input->dim_count; i++) norm *= input->dim[i];
    if (op->t_axis < 0 || op->t_axis >= input->dim_count || !(op->d_gamma = const_as_f32(op, gamma, norm, "layer_norm gamma")) ||
        !(op->d_beta = const_as_f32(op, beta, norm, "layer_norm beta"))) {
        free(op);
        return CSINN_FALSE;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())norm_exec4;
    return CSINN_TRUE;
}
static int rms_norm_exec(struct csinn_tensor *input, struct csinn_tensor *w, struct csinn_tensor *output, void *params)
{
    (void)w;
    return layer_exec(params, input, NULL, output);
}
static int rms_norm_init(struct csinn_tensor *input, struct csinn_tensor *weights, struct csinn_tensor *output,
                         struct csinn_rms_norm_params *params)
{
    b200_op *op = tensor_op_new(&params->base, input, output, B200_T_RMS_NORM, "b200_rms_norm");
    if (!op) return CSINN_FALSE;
    op->t_axis = params->axis >= 0 ? params->axis : params->axis + input->dim_count;
    op->t_eps = params->epsilon;
    int64_t norm = 1;
    for (int i = op->t_axis; i >= 0 && i < input->dim_count; i++) norm *= input->dim[i];
    if (op->t_axis < 0 || op->t_axis >= input->dim_count || !(op->d_gamma = const_as_f32(op, weights, norm, "rms_norm weight"))) {
        free(op);
        return CSINN_FALSE;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())rms_norm_exec;
    return CSINN_TRUE;
}

/* csinn_matmul (source/reference/matmul.c:21; replaces shl_rvv_matmul_int8 / _fp16): out[b][i][j] = sum_k a[b][i][k] *
 * m1[b][k][j] with optional transposes; mat1 shared by all batches when it has none.  A constant mat1 is packed once as
 * [J][K] rows (the fullyconnected weight layout) with the zero-point fold of the requantise contract; an activation mat1
 * (fp16 only: its column sums would be run-time data) is packed per run. */
static int matmul_exec(struct csinn_tensor *mat0, struct csinn_tensor *mat1, struct csinn_tensor *output, void *params)
{
    return layer_exec(params, mat0, mat1->is_const ? NULL : mat1, output);
}
static int matmul_init(struct csinn_tensor *mat0, struct csinn_tensor *mat1, struct csinn_tensor *output,
                       struct csinn_matmul_params *params)
{
    if (mat0->dim_count < 2 || mat1->dim_count < 2 || mat1->dim_count > 4 || mat0->dtype != mat1->dtype) {
        b200_fail("matmul: operands of rank 2..4 and one dtype");
        return CSINN_FALSE;
    }
    b200_op *op = tensor_op_new(&params->base, mat0, output, B200_T_MATMUL, "b200_matmul_tcgen05");
    if (!op) return CSINN_FALSE;
    const int ra = mat0->dim_count, rb = mat1->dim_count;
    op->mm_trans_a = params->trans_a ? 1 : 0, op->mm_trans_b = params->trans_b ? 1 : 0;
    op->mm_i = mat0->dim[ra - (params->trans_a ? 1 : 2)];
    op->mm_k = mat0->dim[ra - (params->trans_a ? 2 : 1)];
    op->mm_j = mat1->dim[rb - (params->trans_b ? 2 : 1)];
    const int kb = mat1->dim[rb - (params->trans_b ? 1 : 2)];
    op->mm_batches = 1, op->mm_batches_b = 1;
    for (int i = 0; i < ra - 2; i++) op->mm_batches *= mat0->dim[i];
    for (int i = 0; i < rb - 2; i++) op->mm_batches_b *= mat1->dim[i];
    op->in1_rank = rb;
    for (int i = 0; i < rb; i++) op->in1_dim[i] = mat1->dim[i];
    op->mm_const_b = mat1->is_const ? 1 : 0;
    op->kdim = op->mm_k, op->ldk = b200_round_channels(op->mm_k, op->eb), op->o = op->mm_j, op->cin = op->mm_k;
    int rc = CSINN_TRUE;
    if (kb != op->mm_k || (op->mm_batches_b != 1 && op->mm_batches_b != op->mm_batches)) {
        b200_fail("matmul: inner extents %d vs %d, batches %d vs %d", op->mm_k, kb, op->mm_batches, op->mm_batches_b);
        rc = CSINN_FALSE;
    } else if (!op->mm_const_b && op->dtype == B200_I8) {
        b200_fail("matmul: int8 needs a constant second operand (its zero-point fold is precomputed)");
        rc = CSINN_FALSE;
    } else if (op->mm_const_b && op->mm_batches_b != 1) {
        b200_fail("matmul: a constant second operand with batches is not supported");
        rc = CSINN_FALSE;
    }
    if (rc == CSINN_TRUE && op->mm_const_b) {
        /* [J][K] copy of mat1, then exactly the fullyconnected machinery: requantise tables + packed rows */
        const int J = op->mm_j, K = op->mm_k, eb = op->eb;
        struct csinn_tensor wt = *mat1;
        uint8_t *rows = malloc((size_t)J * K * eb);
        if (!rows) {
            free(op);
            return CSINN_FALSE;
        }
        const uint8_t *src = mat1->data;
        for (int j = 0; j < J; j++)
            for (int k = 0; k < K; k++)
                memcpy(rows + ((size_t)j * K + k) * eb, src + (params->trans_b ? ((size_t)j * K + k) : ((size_t)k * J + j)) * eb, (size_t)eb);
        wt.data = rows, wt.dim_count = 2, wt.dim[0] = J, wt.dim[1] = K;
        size_t wbytes = 0;
        rc = b200_make_requant(op, mat0, &wt, NULL, output, K, 0, J);
        if (rc == CSINN_TRUE && !(op->d_w = b200_pack_fc_weights(op, &wt, &wbytes))) rc = CSINN_FALSE;
        free(rows);
    } else if (rc == CSINN_TRUE) {
        struct csinn_tensor none = *mat1; /* fp16 x fp16 activations: no tables, no bias */
        none.data = NULL;
        op->d_mult = op->d_badd = NULL, op->d_ibias = NULL;
        op->two_inputs = 1;
        (void)none;
    }
    if (rc != CSINN_TRUE) {
        free(op);
        return CSINN_FALSE;
    }
    b200_op_bind(params, op);
    params->base.cb->exec = (int (*)())matmul_exec;
    return CSINN_TRUE;
}

void *shl_b200_transpose_init_fn(void) { return (void *)transpose_init; }
void *shl_b200_gather_init_fn(void) { return (void *)gather_init; }
void *shl_b200_reduce_sum_init_fn(void) { return (void *)reduce_sum_init; }
void *shl_b200_layer_norm_init_fn(void) { return (void *)layer_norm_init; }
void *shl_b200_rms_norm_init_fn(void) { return (void *)rms_norm_init; }
void *shl_b200_matmul_init_fn(void) { return (void *)matmul_init; }
void *shl_b200_tensor_exec1_fn(void) { return (void *)tensor_exec1; }
void *shl_b200_gather_exec_fn(void) { return (void *)gather_exec; }
void *shl_b200_norm_exec4_fn(void) { return (void *)norm_exec4; }
void *shl_b200_rms_norm_exec_fn(void) { return (void *)rms_norm_exec; }
void *shl_b200_matmul_exec_fn(void) { return (void *)matmul_exec; }
