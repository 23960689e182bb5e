/*
 * setup.c -- registration of the b200 backend in the reference's backend registry.
 *
 * Plays the role of source/thead_rvv/setup.c:68-535 and source/c920_opt/setup.c:24-57,354-386:
 * a key -> csinn_callback{init, est, exec, caps, perf} table (key = op * CSINN_DTYPE_SIZE +
 * dtype, thead_rvv/setup.c:33), an op map and a runtime map, handed to
 * shl_register_op_callback / shl_register_runtime_callback (source/nn2/setup.c:99,129).
 * `est` reuses the reference's own graph recorders shl_gref_<op> exactly as the RVV back end
 * does; `init` / `exec` are ours (ops.c); session hooks are ours (graph.c).
 *
 * shl_target_init_rvv / _c906 / _c908 / _c920 / _c920v2 are the symbols shl_init() calls
 * (source/nn2/setup.c:36-71); defining them here -- the RISC-V directories are not compiled --
 * is how b200 takes those api ids without touching a reference file.
 */
#include <string.h>

#include "b200_internal.h"

#define B200_CB_MAX 160
static struct shl_cb_table g_cb_table[B200_CB_MAX];
static int g_cb_n;
static struct csinn_callback g_cb_unset; /* all NULL: csinn_<op>() -> CSINN_CALLBACK_UNSET */

void *shl_b200_conv2d_relu_init_fn(void);
void *shl_b200_conv2d_relu6_init_fn(void);
void *shl_b200_relu6_init_fn(void);
void *shl_b200_avgpool_init_fn(void);
void *shl_b200_global_avgpool_init_fn(void);

static void reg_op(int dtype, int op, void *init, void *exec, void *est, void *perf)
{
    if (g_cb_n >= B200_CB_MAX) {
        shl_debug_error("b200 callback table is full\n");
        return;
    }
    g_cb_table[g_cb_n].shl_cb_key = op * CSINN_DTYPE_SIZE + dtype;
    g_cb_table[g_cb_n].shl_cb_value.init = init;
    g_cb_table[g_cb_n].shl_cb_value.exec = exec;
    g_cb_table[g_cb_n].shl_cb_value.est = est;
    g_cb_table[g_cb_n].shl_cb_value.caps = NULL;
    g_cb_table[g_cb_n].shl_cb_value.perf = perf;
    g_cb_n++;
}

struct csinn_callback *shl_cb_map_b200(int op, int dtype)
{
    for (int i = 0; i < g_cb_n; i++)
        if (g_cb_table[i].shl_cb_key == op * CSINN_DTYPE_SIZE + dtype)
            return &g_cb_table[i].shl_cb_value;
    /* no fall-through to shl_cb_map_ref (the RVV map does, thead_rvv/setup.c:52-55): there is
     * no CPU path.  nn2 memcpy()s whatever we return (source/nn2/setup.c:118-122), so hand it
     * a zeroed callback rather than NULL. */
    shl_debug_info("b200: op %d dtype %d is not implemented by the b200 backend\n", op, dtype);
    memset(&g_cb_unset, 0, sizeof(g_cb_unset));
    return &g_cb_unset;
}

static void build_table(void)
{
    if (g_cb_n) return;
    const int dts[2] = {CSINN_DTYPE_INT8, CSINN_DTYPE_FLOAT16};
    for (int i = 0; i < 2; i++) {
        const int dt = dts[i];
        reg_op(dt, CSINN_OP_CONV2D, shl_b200_conv2d_init, shl_b200_conv2d, shl_gref_conv2d, shl_b200_perf);
        reg_op(dt, CSINN_OP_GROUP_CONV2D, shl_b200_conv2d_init, shl_b200_conv2d, shl_gref_group_conv2d, shl_b200_perf);
        reg_op(dt, CSINN_OP_GROUP_CONV2D_RELU, shl_b200_conv2d_relu_init_fn(), shl_b200_conv2d, shl_gref_group_conv2d_relu, shl_b200_perf);
        reg_op(dt, CSINN_OP_CONV2D_RELU, shl_b200_conv2d_relu_init_fn(), shl_b200_conv2d, shl_gref_conv2d_relu, shl_b200_perf);
        reg_op(dt, CSINN_OP_CONV2D_RELU6, shl_b200_conv2d_relu6_init_fn(), shl_b200_conv2d, shl_gref_conv2d_relu6, shl_b200_perf);
        reg_op(dt, CSINN_OP_DEPTHWISE_CONV2D, shl_b200_depthwise_conv2d_init, shl_b200_depthwise_conv2d, shl_gref_depthwise_conv2d, shl_b200_perf);
        reg_op(dt, CSINN_OP_DEPTHWISE_CONV2D_RELU, shl_b200_conv2d_relu_init_fn(), shl_b200_depthwise_conv2d, shl_gref_depthwise_conv2d_relu, shl_b200_perf);
        reg_op(dt, CSINN_OP_DEPTHWISE_CONV2D_RELU6, shl_b200_conv2d_relu6_init_fn(), shl_b200_depthwise_conv2d, shl_gref_depthwise_conv2d_relu6, shl_b200_perf);
        reg_op(dt, CSINN_OP_FULLYCONNECTED, shl_b200_fullyconnected_init, shl_b200_fullyconnected, shl_gref_fullyconnected, shl_b200_perf);
        reg_op(dt, CSINN_OP_RELU, shl_b200_relu_init, shl_b200_relu, shl_gref_relu, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_RELU6, shl_b200_relu6_init_fn(), shl_b200_relu, shl_gref_relu6, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_LEAKY_RELU, shl_b200_leaky_relu_init_fn(), shl_b200_relu, shl_gref_leaky_relu, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_SIGMOID, shl_b200_sigmoid_init_fn(), shl_b200_relu, shl_gref_sigmoid, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_SILU, shl_b200_silu_init_fn(), shl_b200_relu, shl_gref_silu, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_ERF, shl_b200_erf_init_fn(), shl_b200_relu, shl_gref_erf, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_CLIP, shl_b200_clip_init_fn(), shl_b200_relu, shl_gref_clip, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_ADD, shl_b200_add_init, shl_b200_add, shl_gref_add, shl_b200_perf_diso);
        reg_op(dt, CSINN_OP_SUB, shl_b200_sub_init_fn(), shl_b200_add, shl_gref_sub, shl_b200_perf_diso);
        reg_op(dt, CSINN_OP_MUL, shl_b200_mul_init_fn(), shl_b200_add, shl_gref_mul, shl_b200_perf_diso);
        reg_op(dt, CSINN_OP_CONCAT, shl_b200_concat_init, shl_b200_concat, shl_gref_concat, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_DIV, shl_b200_div_init_fn(), shl_b200_add, shl_gref_div, shl_b200_perf_diso);
        reg_op(dt, CSINN_OP_PRELU, shl_b200_prelu_init_fn(), shl_b200_add, shl_gref_prelu, shl_b200_perf_diso);
        reg_op(dt, CSINN_OP_SPLIT, shl_b200_split_init, shl_b200_split, shl_gref_split, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_MAXPOOL2D, shl_b200_pool2d_init, shl_b200_pool2d, shl_gref_maxpool2d, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_AVGPOOL2D, shl_b200_avgpool_init_fn(), shl_b200_pool2d, shl_gref_avgpool2d, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_GLOBAL_AVGPOOL2D, shl_b200_global_avgpool_init_fn(), shl_b200_pool2d, shl_gref_global_avgpool2d, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_GLOBAL_MAXPOOL2D, shl_b200_global_maxpool_init_fn(), shl_b200_pool2d, shl_gref_global_maxpool2d, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_SOFTMAX, shl_b200_softmax_init, shl_b200_softmax, shl_gref_softmax, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_RESHAPE, shl_b200_reshape_init, shl_b200_reshape, shl_gref_reshape, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_FLATTEN, shl_b200_reshape_init, shl_b200_reshape, shl_gref_flatten, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_TRANSPOSE, shl_b200_transpose_init_fn(), shl_b200_tensor_exec1_fn(), shl_gref_transpose, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_GATHER, shl_b200_gather_init_fn(), shl_b200_gather_exec_fn(), shl_gref_gather, shl_b200_perf_diso);
        reg_op(dt, CSINN_OP_REDUCE_SUM, shl_b200_reduce_sum_init_fn(), shl_b200_tensor_exec1_fn(), shl_gref_reduce_sum, shl_b200_perf_siso);
        reg_op(dt, CSINN_OP_LAYER_NORM, shl_b200_layer_norm_init_fn(), shl_b200_norm_exec4_fn(), shl_gref_layer_norm, NULL);
        reg_op(dt, CSINN_OP_RMS_NORM, shl_b200_rms_norm_init_fn(), shl_b200_rms_norm_exec_fn(), shl_gref_rms_norm, NULL);
        reg_op(dt, CSINN_OP_MATMUL, shl_b200_matmul_init_fn(), shl_b200_matmul_exec_fn(), shl_gref_matmul, shl_b200_perf_diso);
    }
}

/* ---- runtime map (cf. shl_c920_runtime_callback, source/c920_opt/setup.c:354-386) -------------- */
void *shl_b200_runtime_callback(int api)
{
    switch (api) {
        case CSINN_SESSION_INIT:
            return shl_b200_session_init;
        case CSINN_SESSION_DEINIT:
            return shl_b200_session_deinit;
        case CSINN_SESSION_SETUP:
            return shl_b200_session_setup;
        case CSINN_SESSION_RUN:
            return shl_b200_session_run;
        case CSINN_UPDATE_INPUT:
            return shl_b200_update_input;
        case CSINN_LOAD_BG:
            return shl_b200_load_binary_model;
        case CSINN_UPDATE_OUTPUT:
        case CSINN_SET_INPUT_NUMBER:
        case CSINN_SET_OUTPUT_NUMBER:
        case CSINN_SET_INPUT:
        case CSINN_SET_OUTPUT:
        case CSINN_GET_INPUT:
        case CSINN_GET_OUTPUT:
        case CSINN_TENSOR_ENTRY:
            /* graph recording and I/O bookkeeping are the reference's own */
            return shl_gref_runtime_callback(api);
        default:
            shl_debug_info("%s: no b200 runtime callback for %d\n", __func__, api);
            break;
    }
    return NULL;
}

void shl_target_init_b200(int api)
{
    build_table();
    shl_register_op_callback(api, shl_cb_map_b200);
    shl_register_runtime_callback(api, shl_b200_runtime_callback);
}

void shl_target_init_rvv(void) { shl_target_init_b200(CSINN_RVV); }
void shl_target_init_c906(void) { shl_target_init_b200(CSINN_C906); }
void shl_target_init_c908(void) { shl_target_init_b200(CSINN_C908); }
void shl_target_init_c920(void) { shl_target_init_b200(CSINN_C920); }
void shl_target_init_c920v2(void) { shl_target_init_b200(CSINN_C920V2); }

/* ---- two helpers the reference's tensor-dump code links against -------------------------------
 * source/utils/debug.c:1204-1235 (profiler level DUMP) converts tensors to f32 through
 * shl_ref_tensor_transform_f32 / _free_f32, which live in source/reference/utils.c:526-600.  The
 * reference operator backend is deliberately NOT linked into libshl_b200.so, so the two symbols
 * are provided here on top of the reference's own csinn_tensor_data_convert
 * (source/nn2/utils.c:2206). */
struct csinn_tensor *shl_ref_tensor_transform_f32(struct csinn_tensor *input)
{
    struct csinn_tensor *ret = csinn_alloc_tensor(NULL);
    if (!ret) return NULL;
    ret->dtype = CSINN_DTYPE_FLOAT32;
    ret->layout = input->layout;
    ret->dim_count = input->dim_count;
    memcpy(ret->dim, input->dim, sizeof(ret->dim));
    ret->name = input->name;
    ret->is_const = input->is_const;
    ret->quant_channel = 0;
    ret->qinfo = NULL;
    const int size = csinn_tensor_size(input);
    if (ret->dim_count == 0 || size == 0) return ret;
    ret->data = shl_mem_alloc((int64_t)size * sizeof(float));
    if (!ret->data || csinn_tensor_data_convert(ret, input) != CSINN_TRUE) return NULL;
    return ret;
}

int shl_ref_tensor_transform_free_f32(struct csinn_tensor *input)
{
    if (csinn_tensor_size(input) != 0) shl_mem_free(input->data);
    csinn_free_tensor(input);
    return CSINN_TRUE;
}
