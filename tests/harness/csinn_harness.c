/*
 * csinn_harness.c -- TEST / BENCH INFRASTRUCTURE.  A flat C front end over the CSI-NN2 public
 * API (csinn_alloc_session / csinn_alloc_tensor / csinn_<op>_init / csinn_<op> /
 * csinn_session_setup / csinn_session_run) so that Python can drive the SAME calls a user of the
 * reference makes, through ctypes, against two builds:
 *     libharness_b200.so  -> csi-nn2_b200/lib/libshl_b200.so   (the product, api = CSINN_RVV ...)
 *     libharness_ref.so   -> oracle/_ref/libshl_ref_x86.so     (the unmodified reference, CSINN_REF)
 * It plays the part of tests/validation_layer/testutil.h:845-889 + tests/utils/test_utils.c of
 * the reference (build tensors with qinfo, run the op through the API, hand back raw outputs),
 * written for this repo; nothing of it ships in the product libraries.
 *
 * A network is a list of h_layer; tensor id 0 is the network input, id i+1 the output of layer i.
 * run_mode 0 = CSINN_RM_LAYER (each op executed eagerly with host tensors), 1 = CSINN_RM_CPU_GRAPH.
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csi_nn.h"
#include "shl_utils.h"

enum {
    H_CONV = 0,   /* csinn_conv2d (group decides conv / depthwise / group, nn2/convolution.c:30-37) */
    H_CONV_RELU,  /* csinn_conv2d_relu */
    H_CONV_RELU6, /* csinn_conv2d_relu6 */
    H_DWCONV,     /* csinn_depthwise_conv2d */
    H_FC,
    H_RELU,
    H_RELU6,
    H_ADD,
    H_MAXPOOL,
    H_AVGPOOL,
    H_GAP,
    H_SOFTMAX,
    H_FLATTEN,
    H_RESHAPE,
    H_LEAKY_RELU, /* csinn_leaky_relu, slope in p0 */
    H_SIGMOID,
    H_CLIP,       /* csinn_clip, [p0, p1] */
    H_SUB,
    H_MUL,
    H_CONCAT, /* csinn_concat of (in0, in1) along `axis`; p0 == 3: of (in0, in1, in0) */
    H_SILU,
    H_ERF,
    H_GMP, /* csinn_global_maxpool2d */
    H_PRELU, /* csinn_prelu, slope = the constant operand ([C]) */
    H_SPLIT, /* csinn_split of in0 in two at index (int)p0 along `axis`; the layer's output is slice (int)p1, the
                other slice goes to a scratch tensor (net->k[i]) with the same qinfo */
    H_DIV,
    H_TRANSPOSE,  /* csinn_transpose, permutation in (kh, kw, sh, sw) */
    H_GATHER,     /* csinn_gather along `axis`, `o` constant int64 indices in w */
    H_REDUCE_SUM, /* csinn_reduce_sum over `axis` (-1: everything); keepdims when the ranks agree */
    H_LAYER_NORM, /* csinn_layer_norm from `axis` on, eps = p0, gamma = w, beta = b (o elements, qinfo s_w / zp_w) */
    H_RMS_NORM,   /* csinn_rms_norm from `axis` on, eps = p0, weight = w */
    H_MATMUL,     /* csinn_matmul(in0, mat1): mat1 = constant w with the pt dims (kh, kw, sh, sw), or tensor in1 when w is
                     NULL; trans_a = pd, trans_b = pr */
};

typedef struct {
    int32_t kind;
    int32_t in0, in1;
    int32_t out_dims[4];
    int32_t out_rank;
    float s_out;
    int32_t zp_out;
    int32_t o, kh, kw, sh, sw, pt, pl, pd, pr, dh, dw, group, fuse_zp2bias;
    const void *w;
    const void *b;
    const float *s_w;
    const int32_t *zp_w;
    int32_t w_channels;
    const float *s_b;
    int32_t count_include_pad, ceil_mode, axis;
    float p0, p1; /* unary-op parameters */
    int32_t w_int8; /* fp16 network: this layer's weights are int8 with qinfo s_w / zp_w (CSINN_QUANT_FLOAT16_W_INT8) */
} h_layer;

typedef struct {
    int api, dtype, run_mode, n;
    struct csinn_session *sess;
    struct csinn_tensor **t; /* n + 1 activation tensors */
    struct csinn_tensor **k, **bias;
    void **params;
    h_layer *layers;
    int setup_done;
} h_net;

#ifdef HARNESS_B200
int shl_b200_session_prefetch_input(int index, const void *host_ptr, struct csinn_session *sess);
void shl_b200_op_release(void *params);
int shl_b200_error_count(void);
const char *shl_b200_last_error(void);
#define BACKEND_ERRORS() shl_b200_error_count()
#else
#define BACKEND_ERRORS() 0
#endif

static char g_err[512];
static void set_err(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *h_last_error(void) { return g_err; }

static int elem_bytes(int dtype)
{
    switch (dtype) {
        case CSINN_DTYPE_INT8:
        case CSINN_DTYPE_UINT8:
            return 1;
        case CSINN_DTYPE_FLOAT16:
            return 2;
        default:
            return 4;
    }
}

static struct csinn_tensor *new_tensor(h_net *net, const char *name, const int32_t *dims, int rank,
                                       int dtype, int layout, int is_const, int qch)
{
    struct csinn_tensor *t = csinn_alloc_tensor(net->sess);
    t->name = strdup(name);
    t->dtype = dtype;
    t->layout = layout;
    t->dim_count = rank;
    for (int i = 0; i < rank; i++) t->dim[i] = dims[i];
    t->is_const = is_const;
    t->quant_channel = qch;
    if (qch > 1) {
        /* csinn_alloc_tensor gives one qinfo (nn2/utils.c:382); per-channel needs an array */
        t->qinfo = calloc(qch, sizeof(struct csinn_quant_info));
    }
    for (int i = 0; i < (qch > 0 ? qch : 1); i++) {
        t->qinfo[i].scale = 1.0f;
        t->qinfo[i].zero_point = 0;
    }
    return t;
}

static int act_layout(int rank)
{
    switch (rank) {
        case 4:
            return CSINN_LAYOUT_NCHW;
        case 3:
            return CSINN_LAYOUT_NCW;
        case 2:
            return CSINN_LAYOUT_NC;
        default:
            return CSINN_LAYOUT_N;
    }
}

static int64_t tsize(const struct csinn_tensor *t)
{
    int64_t s = 1;
    for (int i = 0; i < t->dim_count; i++) s *= t->dim[i];
    return t->dim_count ? s : 0;
}

static void base_init(h_net *net, struct csinn_params_base *b, const char *name)
{
    b->name = strdup(name);
    b->layout = CSINN_LAYOUT_NCHW;
    b->api = net->api;
    b->quant_type = net->dtype == CSINN_DTYPE_INT8
                        ? CSINN_QUANT_INT8_ASYM_W_SYM
                        : (net->dtype == CSINN_DTYPE_FLOAT16 ? CSINN_QUANT_FLOAT16 : CSINN_QUANT_FLOAT32);
}

static int layer_init(h_net *net, int i)
{
    h_layer *L = &net->layers[i];
    struct csinn_tensor *in = net->t[L->in0], *out = net->t[i + 1];
    char nm[64];
    snprintf(nm, sizeof(nm), "layer_%d", i);
    switch (L->kind) {
        case H_CONV:
        case H_CONV_RELU:
        case H_CONV_RELU6:
        case H_DWCONV: {
            struct csinn_conv2d_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->group = L->group, p->stride_height = L->sh, p->stride_width = L->sw;
            p->pad_top = L->pt, p->pad_left = L->pl, p->pad_down = L->pd, p->pad_right = L->pr;
            p->dilation_height = L->dh, p->dilation_width = L->dw;
            p->conv_extra.kernel_tm = NULL, p->conv_extra.conv_mode = CSINN_DIRECT;
            p->conv_extra.fuse_zp2bias = L->fuse_zp2bias;
            net->params[i] = p;
            switch (L->kind) {
                case H_CONV:
                    return csinn_conv2d_init(in, out, net->k[i], net->bias[i], p);
                case H_CONV_RELU:
                    return csinn_conv2d_relu_init(in, out, net->k[i], net->bias[i], p);
                case H_CONV_RELU6:
                    return csinn_conv2d_relu6_init(in, out, net->k[i], net->bias[i], p);
                default:
                    return csinn_depthwise_conv2d_init(in, out, net->k[i], net->bias[i], p);
            }
        }
        case H_FC: {
            struct csinn_fc_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->units = L->o;
            p->fc_extra.fuse_zp2bias = L->fuse_zp2bias;
            net->params[i] = p;
            return csinn_fullyconnected_init(in, out, net->k[i], net->bias[i], p);
        }
        case H_RELU:
        case H_RELU6: {
            struct csinn_relu_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            net->params[i] = p;
            return L->kind == H_RELU ? csinn_relu_init(in, out, p) : csinn_relu6_init(in, out, p);
        }
        case H_LEAKY_RELU: {
            struct csinn_relu_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->n = L->p0;
            net->params[i] = p;
            return csinn_leaky_relu_init(in, out, p);
        }
        case H_SIGMOID: {
            struct csinn_sigmoid_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            net->params[i] = p;
            return csinn_sigmoid_init(in, out, p);
        }
        case H_SILU: {
            struct csinn_sigmoid_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            net->params[i] = p;
            return csinn_silu_init(in, out, p);
        }
        case H_ERF: {
            struct csinn_siso_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            net->params[i] = p;
            return csinn_erf_init(in, out, p);
        }
        case H_CLIP: {
            struct csinn_clip_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->min_value = L->p0, p->max_value = L->p1;
            net->params[i] = p;
            return csinn_clip_init(in, out, p);
        }
        case H_ADD:
        case H_SUB:
        case H_DIV:
        case H_MUL: {
            struct csinn_diso_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            net->params[i] = p;
            struct csinn_tensor *rhs = L->w ? net->k[i] : net->t[L->in1];
            if (L->kind == H_SUB) return csinn_sub_init(in, rhs, out, p);
            if (L->kind == H_DIV) return csinn_div_init(in, rhs, out, p);
            if (L->kind == H_MUL) return csinn_mul_init(in, rhs, out, p);
            return csinn_add_init(in, rhs, out, p);
        }
        case H_SPLIT: {
            struct csinn_split_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->output_num = 2, p->axis = L->axis;
            p->split_index = calloc(2, sizeof(int32_t));
            p->split_index[0] = (int32_t)L->p0, p->split_index[1] = in->dim[L->axis];
            net->params[i] = p;
            struct csinn_tensor *outs[2] = {L->p1 != 0.f ? net->k[i] : out, L->p1 != 0.f ? out : net->k[i]};
            return csinn_split_init(in, outs, p);
        }
        case H_PRELU: {
            struct csinn_prelu_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->axis = 1;
            net->params[i] = p;
            return csinn_prelu_init(in, net->k[i], out, p);
        }
        case H_TRANSPOSE: {
            struct csinn_transpose_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->permute_num = in->dim_count;
            p->permute = calloc(4, sizeof(int32_t));
            p->permute[0] = L->kh, p->permute[1] = L->kw, p->permute[2] = L->sh, p->permute[3] = L->sw;
            net->params[i] = p;
            return csinn_transpose_init(in, out, p);
        }
        case H_GATHER: {
            struct csinn_gather_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->axis = L->axis;
            net->params[i] = p;
            return csinn_gather_init(in, net->k[i], out, p);
        }
        case H_REDUCE_SUM: {
            struct csinn_reduce_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->axis_count = 1;
            p->axis = calloc(1, sizeof(int32_t));
            p->axis[0] = L->axis;
            p->keepdims = out->dim_count == in->dim_count;
            net->params[i] = p;
            return csinn_reduce_sum_init(in, out, p);
        }
        case H_LAYER_NORM: {
            struct csinn_layer_norm_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->epsilon = L->p0, p->axis = L->axis, p->center = true, p->scale = true;
            net->params[i] = p;
            return csinn_layer_norm_init(in, out, net->k[i], net->bias[i], p);
        }
        case H_RMS_NORM: {
            struct csinn_rms_norm_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->epsilon = L->p0, p->axis = L->axis;
            net->params[i] = p;
            return csinn_rms_norm_init(in, net->k[i], out, p);
        }
        case H_MATMUL: {
            struct csinn_matmul_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->trans_a = L->pd != 0, p->trans_b = L->pr != 0;
            net->params[i] = p;
            return csinn_matmul_init(in, L->w ? net->k[i] : net->t[L->in1], out, p);
        }
        case H_CONCAT: {
            struct csinn_concat_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->inputs_count = L->p0 == 3.f ? 3 : 2, p->axis = L->axis;
            net->params[i] = p;
            struct csinn_tensor *ins[3] = {in, net->t[L->in1], in};
            return csinn_concat_init(ins, out, p);
        }
        case H_MAXPOOL:
        case H_AVGPOOL:
        case H_GMP:
        case H_GAP: {
            struct csinn_pool_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->filter_height = L->kh, p->filter_width = L->kw;
            p->stride_height = L->sh, p->stride_width = L->sw;
            p->pad_top = L->pt, p->pad_left = L->pl, p->pad_down = L->pd, p->pad_right = L->pr;
            p->count_include_pad = L->count_include_pad, p->ceil_mode = L->ceil_mode;
            net->params[i] = p;
            if (L->kind == H_MAXPOOL) return csinn_maxpool2d_init(in, out, p);
            if (L->kind == H_AVGPOOL) return csinn_avgpool2d_init(in, out, p);
            if (L->kind == H_GMP) return csinn_global_maxpool2d_init(in, out, p);
            return csinn_global_avgpool2d_init(in, out, p);
        }
        case H_SOFTMAX: {
            struct csinn_softmax_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->axis = L->axis;
            net->params[i] = p;
            return csinn_softmax_init(in, out, p);
        }
        case H_FLATTEN: {
            struct csinn_flatten_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->axis = L->axis;
            net->params[i] = p;
            return csinn_flatten_init(in, out, p);
        }
        case H_RESHAPE: {
            struct csinn_reshape_params *p = csinn_alloc_params(sizeof(*p), net->sess);
            base_init(net, &p->base, nm);
            p->shape = L->out_dims, p->shape_num = L->out_rank;
            net->params[i] = p;
            return csinn_reshape_init(in, out, p);
        }
    }
    set_err("layer %d: unknown kind %d", i, L->kind);
    return CSINN_FALSE;
}

static int layer_call(h_net *net, int i)
{
    h_layer *L = &net->layers[i];
    struct csinn_tensor *in = net->t[L->in0], *out = net->t[i + 1];
    void *p = net->params[i];
    switch (L->kind) {
        case H_CONV:
            return csinn_conv2d(in, out, net->k[i], net->bias[i], p);
        case H_CONV_RELU:
            return csinn_conv2d_relu(in, out, net->k[i], net->bias[i], p);
        case H_CONV_RELU6:
            return csinn_conv2d_relu6(in, out, net->k[i], net->bias[i], p);
        case H_DWCONV:
            return csinn_depthwise_conv2d(in, out, net->k[i], net->bias[i], p);
        case H_FC:
            return csinn_fullyconnected(in, out, net->k[i], net->bias[i], p);
        case H_RELU:
            return csinn_relu(in, out, p);
        case H_RELU6:
            return csinn_relu6(in, out, p);
        case H_LEAKY_RELU:
            return csinn_leaky_relu(in, out, p);
        case H_SIGMOID:
            return csinn_sigmoid(in, out, p);
        case H_CLIP:
            return csinn_clip(in, out, p);
        case H_SILU:
            return csinn_silu(in, out, p);
        case H_ERF:
            return csinn_erf(in, out, p);
        case H_SUB:
            return csinn_sub(in, L->w ? net->k[i] : net->t[L->in1], out, p);
        case H_DIV:
            return csinn_div(in, L->w ? net->k[i] : net->t[L->in1], out, p);
        case H_MUL:
            return csinn_mul(in, L->w ? net->k[i] : net->t[L->in1], out, p);
        case H_ADD:
            return csinn_add(in, L->w ? net->k[i] : net->t[L->in1], out, p);
        case H_SPLIT: {
            struct csinn_tensor *outs[2] = {L->p1 != 0.f ? net->k[i] : out, L->p1 != 0.f ? out : net->k[i]};
            return csinn_split(in, outs, p);
        }
        case H_PRELU:
            return csinn_prelu(in, net->k[i], out, p);
        case H_TRANSPOSE:
            return csinn_transpose(in, out, p);
        case H_GATHER:
            return csinn_gather(in, net->k[i], out, p);
        case H_REDUCE_SUM:
            return csinn_reduce_sum(in, out, p);
        case H_LAYER_NORM:
            return csinn_layer_norm(in, out, net->k[i], net->bias[i], p);
        case H_RMS_NORM:
            return csinn_rms_norm(in, net->k[i], out, p);
        case H_MATMUL:
            return csinn_matmul(in, L->w ? net->k[i] : net->t[L->in1], out, p);
        case H_CONCAT: {
            struct csinn_tensor *ins[3] = {in, net->t[L->in1], in};
            return csinn_concat(ins, out, p);
        }
        case H_MAXPOOL:
            return csinn_maxpool2d(in, out, p);
        case H_AVGPOOL:
            return csinn_avgpool2d(in, out, p);
        case H_GAP:
            return csinn_global_avgpool2d(in, out, p);
        case H_GMP:
            return csinn_global_maxpool2d(in, out, p);
        case H_SOFTMAX:
            return csinn_softmax(in, out, p);
        case H_FLATTEN:
            return csinn_flatten(in, out, p);
        case H_RESHAPE:
            return csinn_reshape(in, out, p);
    }
    return CSINN_FALSE;
}

void h_net_destroy(void *handle);

/* the next graph-mode h_net_create saves the model in the HHB binary format at session_setup
 * (sess->model.save_mode = CSINN_SAVE_AND_RUN, as a model.c generated with `--save` would) */
static char g_save_path[512];
void h_set_save_path(const char *path) { snprintf(g_save_path, sizeof(g_save_path), "%s", path ? path : ""); }

void *h_net_create(int api, int dtype, int run_mode, const int32_t *in_dims, int in_rank, float s_in,
                   int zp_in, const h_layer *layers, int n)
{
    g_err[0] = 0;
    h_net *net = calloc(1, sizeof(*net));
    net->api = api, net->dtype = dtype, net->run_mode = run_mode, net->n = n;
    net->layers = malloc(sizeof(h_layer) * n);
    memcpy(net->layers, layers, sizeof(h_layer) * n);
    net->t = calloc(n + 1, sizeof(void *));
    net->k = calloc(n, sizeof(void *));
    net->bias = calloc(n, sizeof(void *));
    net->params = calloc(n, sizeof(void *));

    struct csinn_session *sess = csinn_alloc_session();
    net->sess = sess;
    sess->base_api = api;
    sess->base_dtype = dtype;
    sess->base_layout = CSINN_LAYOUT_NCHW;
    sess->base_quant_type = dtype == CSINN_DTYPE_INT8 ? CSINN_QUANT_INT8_ASYM_W_SYM : CSINN_QUANT_UNSET;
    sess->debug_level = CSINN_DEBUG_LEVEL_WARNING;
    if (run_mode == 0) {
        sess->base_run_mode = CSINN_RM_LAYER;
    } else {
        sess->base_run_mode = CSINN_RM_CPU_GRAPH;
        csinn_session_init(sess);
        csinn_set_input_number(1, sess);
        csinn_set_output_number(1, sess);
    }

    net->t[0] = new_tensor(net, "input", in_dims, in_rank, dtype, act_layout(in_rank), 0, 1);
    net->t[0]->qinfo->scale = s_in, net->t[0]->qinfo->zero_point = zp_in;
    const int wdtype = dtype;
    const int bdtype = dtype == CSINN_DTYPE_INT8 ? CSINN_DTYPE_INT32 : dtype;
    for (int i = 0; i < n; i++) {
        const h_layer *L = &net->layers[i];
        char nm[64];
        snprintf(nm, sizeof(nm), "output_%d", i);
        net->t[i + 1] = new_tensor(net, nm, L->out_dims, L->out_rank, dtype, act_layout(L->out_rank), 0, 1);
        net->t[i + 1]->qinfo->scale = L->s_out, net->t[i + 1]->qinfo->zero_point = L->zp_out;
        if (L->kind == H_SPLIT) {
            const struct csinn_tensor *src = net->t[L->in0];
            int32_t od[4];
            for (int d = 0; d < src->dim_count; d++) od[d] = src->dim[d];
            od[L->axis] = src->dim[L->axis] - L->out_dims[L->axis];
            snprintf(nm, sizeof(nm), "split_other_%d", i);
            net->k[i] = new_tensor(net, nm, od, src->dim_count, dtype, act_layout(src->dim_count), 0, 1);
            net->k[i]->qinfo->scale = L->s_out, net->k[i]->qinfo->zero_point = L->zp_out;
            net->k[i]->data = calloc(1, tsize(net->k[i]) * elem_bytes(dtype) + 64);
        }
        if (L->kind == H_PRELU && L->w) {
            int32_t cd[1] = {L->o};
            snprintf(nm, sizeof(nm), "alpha_%d", i);
            net->k[i] = new_tensor(net, nm, cd, 1, wdtype, CSINN_LAYOUT_O, 1, 1);
            net->k[i]->data = (void *)L->w;
            net->k[i]->mtype = CSINN_MEM_TYPE_CPU_ALIGNED;
            net->k[i]->qinfo->scale = L->s_w ? L->s_w[0] : 1.0f;
            net->k[i]->qinfo->zero_point = L->zp_w ? L->zp_w[0] : 0;
        }
        if ((L->kind == H_ADD || L->kind == H_SUB || L->kind == H_MUL || L->kind == H_DIV) && L->w) {
            /* constant second operand: one element ([1]) or one value per channel ([1, C, 1, 1]) */
            int32_t cd[4] = {1, L->o, 1, 1};
            snprintf(nm, sizeof(nm), "const_%d", i);
            net->k[i] = L->o == 1 ? new_tensor(net, nm, cd, 1, wdtype, CSINN_LAYOUT_O, 1, 1)
                                  : new_tensor(net, nm, cd, 4, wdtype, CSINN_LAYOUT_NCHW, 1, 1);
            net->k[i]->data = (void *)L->w;
            net->k[i]->mtype = CSINN_MEM_TYPE_CPU_ALIGNED;
            net->k[i]->qinfo->scale = L->s_w ? L->s_w[0] : 1.0f;
            net->k[i]->qinfo->zero_point = L->zp_w ? L->zp_w[0] : 0;
        }
        if ((L->kind == H_LAYER_NORM || L->kind == H_RMS_NORM) && L->w) {
            int32_t cd[1] = {L->o};
            snprintf(nm, sizeof(nm), "gamma_%d", i);
            net->k[i] = new_tensor(net, nm, cd, 1, wdtype, CSINN_LAYOUT_O, 1, 1);
            net->k[i]->data = (void *)L->w, net->k[i]->mtype = CSINN_MEM_TYPE_CPU_ALIGNED;
            net->k[i]->qinfo->scale = L->s_w ? L->s_w[0] : 1.0f, net->k[i]->qinfo->zero_point = L->zp_w ? L->zp_w[0] : 0;
            if (L->b) {
                snprintf(nm, sizeof(nm), "beta_%d", i);
                net->bias[i] = new_tensor(net, nm, cd, 1, wdtype, CSINN_LAYOUT_O, 1, 1);
                net->bias[i]->data = (void *)L->b, net->bias[i]->mtype = CSINN_MEM_TYPE_CPU_ALIGNED;
                net->bias[i]->qinfo->scale = net->k[i]->qinfo->scale, net->bias[i]->qinfo->zero_point = net->k[i]->qinfo->zero_point;
            }
        }
        if (L->kind == H_GATHER) {
            int32_t cd[1] = {L->o};
            snprintf(nm, sizeof(nm), "indices_%d", i);
            net->k[i] = new_tensor(net, nm, cd, 1, CSINN_DTYPE_INT64, CSINN_LAYOUT_N, 1, 1);
            net->k[i]->data = (void *)L->w, net->k[i]->mtype = CSINN_MEM_TYPE_CPU_ALIGNED;
        }
        if (L->kind == H_MATMUL && L->w) {
            int32_t cd[4] = {L->kh, L->kw, L->sh, L->sw};
            snprintf(nm, sizeof(nm), "mat1_%d", i);
            net->k[i] = new_tensor(net, nm, cd, L->pt, wdtype, act_layout(L->pt), 1, 1);
            net->k[i]->data = (void *)L->w, net->k[i]->mtype = CSINN_MEM_TYPE_CPU_ALIGNED;
            net->k[i]->qinfo->scale = L->s_w ? L->s_w[0] : 1.0f, net->k[i]->qinfo->zero_point = L->zp_w ? L->zp_w[0] : 0;
        }
        if (L->kind <= H_FC) {
            struct csinn_tensor *in = net->t[L->in0];
            int32_t kd[4];
            int krank, klayout;
            if (L->kind == H_FC) {
                kd[0] = L->o, kd[1] = in->dim[in->dim_count - 1], krank = 2, klayout = CSINN_LAYOUT_OI;
                if (in->dim_count == 4) kd[1] = in->dim[1] * in->dim[2] * in->dim[3];
            } else {
                kd[0] = L->o, kd[1] = in->dim[1] / (L->group > 0 ? L->group : 1), kd[2] = L->kh, kd[3] = L->kw;
                krank = 4;
                klayout = (L->group == in->dim[1] && kd[1] == 1 && L->group > 1) ? CSINN_LAYOUT_O1HW : CSINN_LAYOUT_OIHW;
            }
            const int qch = L->w_channels > 0 ? L->w_channels : 1;
            snprintf(nm, sizeof(nm), "kernel_%d", i);
            net->k[i] = new_tensor(net, nm, kd, krank, L->w_int8 ? CSINN_DTYPE_INT8 : wdtype, klayout, 1, qch);
            net->k[i]->data = (void *)L->w;
            net->k[i]->mtype = CSINN_MEM_TYPE_CPU_ALIGNED;
            int32_t bd[1] = {L->o};
            snprintf(nm, sizeof(nm), "bias_%d", i);
            /* an fp16 bias under int8 weights carries one qinfo with scale 1: the reference multiplies fp16
             * constants by qinfo->scale when it is not 1 (f16_to_float, source/nn2/utils.c:1175-1189) */
            const int bqch = L->w_int8 ? 1 : qch;
            net->bias[i] = new_tensor(net, nm, bd, L->b ? 1 : 0, bdtype, CSINN_LAYOUT_O, 1, bqch);
            net->bias[i]->data = (void *)L->b;
            for (int c = 0; c < qch; c++) {
                net->k[i]->qinfo[c].scale = L->s_w ? L->s_w[c] : 1.0f;
                net->k[i]->qinfo[c].zero_point = L->zp_w ? L->zp_w[c] : 0;
                if (c >= bqch) continue;
                net->bias[i]->qinfo[c].scale =
                    L->w_int8 ? 1.0f : (L->s_b ? L->s_b[c] : in->qinfo->scale * net->k[i]->qinfo[c].scale);
                net->bias[i]->qinfo[c].zero_point = 0;
            }
        }
    }

    /* same order as an HHB-generated model.c (example/c906_mobilenetv1_f16.c:22-1960):
     * all *_init, tensor entry + input, all op calls, output, session_setup */
    const int errs0 = BACKEND_ERRORS();
    for (int i = 0; i < n; i++) {
        int rc = layer_init(net, i);
        /* the nn2 front ends discard the callback's status: ask the backend itself */
        if (rc == CSINN_TRUE && BACKEND_ERRORS() != errs0) rc = CSINN_FALSE;
        if (rc != CSINN_TRUE) {
            if (!g_err[0]) set_err("layer %d (kind %d): csinn_*_init returned %d", i, layers[i].kind, rc);
            h_net_destroy(net);
            return NULL;
        }
    }
    if (run_mode != 0) {
        csinn_set_tensor_entry(net->t[0], sess);
        csinn_set_input(0, net->t[0], sess);
        for (int i = 0; i < n; i++) {
            int rc = layer_call(net, i);
            if (rc != CSINN_TRUE) {
                set_err("layer %d (kind %d): graph recording returned %d", i, layers[i].kind, rc);
                h_net_destroy(net);
                return NULL;
            }
        }
        csinn_set_output(0, net->t[n], sess);
        /* a zero-initialised session means CSINN_SAVE_AND_RUN (= 0) and would write shl.hhb.bm into the
         * cwd at every session_setup: only save when a test asked for it */
        sess->model.save_mode = CSINN_RUN_ONLY;
        if (g_save_path[0]) {
            sess->model.save_mode = CSINN_SAVE_AND_RUN;
            sess->model.bm_path = strdup(g_save_path);
            g_save_path[0] = 0;
        }
        int rc = csinn_session_setup(sess);
        net->setup_done = 1;
        /* the reference's own session_setup hooks return void (graph_ref/setup.c:688), so only a
         * definite CSINN_FALSE from a backend that reports status counts as failure */
        if (rc == CSINN_FALSE && api != CSINN_REF) {
            set_err("csinn_session_setup failed");
            h_net_destroy(net);
            return NULL;
        }
    }
    return net;
}

long long h_net_output_bytes(void *handle)
{
    h_net *net = handle;
    return tsize(net->t[net->n]) * elem_bytes(net->dtype);
}

/* a session restored from the HHB binary format, as an HHB-generated main would do it:
 * csinn_import_binary_model(blob) and then update_input / session_run / get_output.  The blob is
 * copied (the loader fixes offsets up in place and the session keeps pointing into it). */
void *h_net_import(const void *blob, long long size)
{
    g_err[0] = 0;
    char *copy = malloc((size_t)size);
    memcpy(copy, blob, (size_t)size);
    const int errs0 = BACKEND_ERRORS();
    struct csinn_session *sess = csinn_import_binary_model(copy);
    if (!sess || BACKEND_ERRORS() != errs0) {
        set_err("csinn_import_binary_model failed");
        return NULL;
    }
    h_net *net = calloc(1, sizeof(*net));
    net->api = sess->base_api, net->dtype = sess->base_dtype, net->run_mode = sess->base_run_mode, net->n = 0;
    net->sess = sess;
    net->t = calloc(1, sizeof(void *));
    net->t[0] = sess->output[0]; /* h_net_output_bytes: dims of the graph output */
    net->setup_done = 1;
    return net;
}

void *h_net_session(void *handle) { return ((h_net *)handle)->sess; }
void *h_net_params(void *handle, int i) { return ((h_net *)handle)->params[i]; }

int h_net_run(void *handle, const void *input, void *output)
{
    h_net *net = handle;
    const int n = net->n;
    if (net->run_mode == 0) {
        void **bufs = calloc(n + 1, sizeof(void *));
        net->t[0]->data = (void *)input;
        int rc = CSINN_TRUE;
        for (int i = 0; i < n && rc == CSINN_TRUE; i++) {
            bufs[i + 1] = i + 1 == n ? output : calloc(1, tsize(net->t[i + 1]) * elem_bytes(net->dtype) + 64);
            net->t[i + 1]->data = bufs[i + 1];
            const int errs0 = BACKEND_ERRORS();
            rc = layer_call(net, i);
            if (rc == CSINN_TRUE && BACKEND_ERRORS() != errs0) rc = CSINN_FALSE;
            if (rc != CSINN_TRUE) set_err("layer %d (kind %d): csinn op returned %d", i, net->layers[i].kind, rc);
        }
        for (int i = 1; i < n; i++) free(bufs[i]);
        free(bufs);
        return rc == CSINN_TRUE ? 0 : -1;
    }
    struct csinn_tensor in_t;
    memset(&in_t, 0, sizeof(in_t));
    in_t.data = (void *)input;
    csinn_update_input(0, &in_t, net->sess);
    int rc = csinn_session_run(net->sess);
    if (rc != CSINN_TRUE) {
        set_err("csinn_session_run returned %d", rc);
        return -1;
    }
    struct csinn_tensor *out_t = csinn_alloc_tensor(NULL);
    csinn_get_output(0, out_t, net->sess);
    if (!out_t->data) {
        set_err("csinn_get_output: no data");
        return -1;
    }
    memcpy(output, out_t->data, h_net_output_bytes(net));
    if (net->api == CSINN_REF) shl_mem_free(out_t->data); /* gref allocates outputs per run (graph_ref/setup.c:1125) */
    csinn_free_tensor(out_t);
    return 0;
}

/* split phases for benchmarking the graph path */
int h_net_update_input(void *handle, const void *input)
{
    h_net *net = handle;
    struct csinn_tensor in_t;
    memset(&in_t, 0, sizeof(in_t));
    in_t.data = (void *)input;
    /* gref's update_input hook returns void (graph_ref/setup.c:51): nn2 hands back garbage */
    csinn_update_input(0, &in_t, net->sess);
    return 0;
}
#ifdef HARNESS_B200
int h_net_prefetch_input(void *handle, const void *input)
{
    return shl_b200_session_prefetch_input(0, input, ((h_net *)handle)->sess) == CSINN_TRUE ? 0 : -1;
}
#endif
int h_net_session_run(void *handle) { return csinn_session_run(((h_net *)handle)->sess) == CSINN_TRUE ? 0 : -1; }
const void *h_net_get_output(void *handle)
{
    h_net *net = handle;
    struct csinn_tensor *out_t = csinn_alloc_tensor(NULL);
    csinn_get_output(0, out_t, net->sess);
    const void *p = out_t->data;
    csinn_free_tensor(out_t);
    return p;
}


void h_net_destroy(void *handle)
{
    h_net *net = handle;
    if (!net) return;
#ifdef HARNESS_B200
    if (net->run_mode == 0)
        for (int i = 0; i < net->n; i++)
            if (net->params[i]) shl_b200_op_release(net->params[i]);
#endif
    if (net->run_mode != 0 && net->sess) csinn_session_deinit(net->sess);
    if (net->sess) csinn_free_session(net->sess);
    free(net->t);
    free(net->k);
    free(net->bias);
    free(net->params);
    free(net->layers);
    free(net);
}

int h_layer_sizeof(void) { return (int)sizeof(h_layer); }
