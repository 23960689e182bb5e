/*
 * b200_internal.h -- private types of the b200 backend (host side, C).
 *
 * Layering:  csinn_* API (reference, unchanged)  ->  callbacks in ops.c / session hooks in
 * graph.c  ->  b200_op (one device operator: packed weights + tables in the weight arena, a
 * run function)  ->  include/b200nn.h (CUDA shim).  The same b200_op serves layer mode
 * (H2D -> convert -> run -> convert -> D2H per call) and graph mode (planned once, replayed
 * as a CUDA graph).
 */
#ifndef B200_INTERNAL_H_
#define B200_INTERNAL_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "b200nn.h"
#include "csi_nn.h"
#include "shl_b200.h"
#include "shl_gref.h"
#include "shl_utils.h"

/* ---- device context: one per session (graph mode) + one process-wide for layer mode ------ */
typedef struct b200_ctx {
    int device;
    void *stream;
    void *copy_stream; /* input prefetch (graph.c), created on first use */
    /* weight arena: bump allocator over one device allocation (graph mode) or a chain of
     * chunks (layer mode) */
    uint8_t *wbase;
    size_t wcap, wused;
    int fixed_arena; /* 1: allocation failure instead of chaining a new chunk */
    int skip_upload; /* 1: allocate but do not copy (weights arrive by NCCL broadcast) */
} b200_ctx;

b200_ctx *b200_ctx_default(void);                        /* NULL + error if no device */
b200_ctx *b200_ctx_of(struct csinn_session *sess);       /* session ctx or the default */
int b200_ctx_init(b200_ctx *ctx, int device);
void b200_ctx_destroy(b200_ctx *ctx);
/* device memory for constants, 256-byte aligned; uploads `src` when non-NULL */
void *b200_warena_put(b200_ctx *ctx, const void *src, size_t bytes);
void b200_fail(const char *fmt, ...);
int b200_default_device(void);

/* ---- device tensor (pixel-major [n][h][w][cp], or raw NCHW for a graph input) ------------- */
typedef struct b200_dt {
    void *d;
    int n, c, h, w, cp;
    int eb;      /* element bytes */
    int is_nchw; /* 1: raw API layout [n][c][h][w], cp unused */
} b200_dt;
size_t b200_dt_bytes(const b200_dt *t);
/* shape of a csinn tensor as the device sees it; returns 0 on unsupported rank */
int b200_dt_from_tensor(b200_dt *t, const struct csinn_tensor *src);

/* ---- one device operator ------------------------------------------------------------------ */
enum b200_op_kind {
    B200_OPK_CONV = 1, /* 1x1 direct GEMM or im2col + GEMM, groups */
    B200_OPK_DW,
    B200_OPK_FC,
    B200_OPK_ACT,     /* relu / relu6 / requantising identity */
    B200_OPK_ADD,
    B200_OPK_POOL,
    B200_OPK_SOFTMAX,
    B200_OPK_COPY,    /* reshape / flatten whose memory order is unchanged */
    B200_OPK_CONCAT,  /* one step per input: its slice of the output (b200_op_run's `part`) */
    B200_OPK_SPLIT,   /* one step per output: its slice of the input (`part` = output index; cat_* fields) */
    B200_OPK_TENSOR,  /* transpose / gather / reduce_sum / layer_norm / rms_norm / matmul (t_* fields, csrc/tensor_ops.cu) */
};

#define B200_CONCAT_MAX 32
typedef struct b200_op {
    int kind;
    int dtype; /* b200_dtype */
    int eb;
    const char *kname; /* kernel name reported to the trace profiler */
    b200_ctx *ctx;
    /* conv / dw / fc */
    int cin, o, kh, kw, sh, sw, pt, pl, dh, dw, group;
    int direct;  /* 1x1 stride-1 unpadded conv or fc: GEMM straight on the activation */
    int kdim;    /* reduction length per group */
    int ldk;     /* im2col / weight row pitch (elements) */
    void *d_w;   /* packed weights */
    void *d_w2;  /* second packing of the same weights (int8 3x3 depthwise: ky-major dp4a words) */
    float *d_mult, *d_badd;
    int32_t *d_ibias;
    /* implicit-GEMM convolution (csrc/gemm_tc.cu IGEMM): border classes of the output positions and the
     * accumulator seeds per (class, output channel); ig_ncls == 0: the op takes the explicit im2col path */
    int ig_ncls;
    void *d_w_diag;  /* depthwise as an implicit GEMM: [cp][taps * 64] rows, diagonal per tap (quant.c b200_pack_dw_diag) */
    int ldk_diag;
    int32_t *d_ig_seeds;
    uint8_t *d_ig_clsmap;
    int ig_h, ig_w, ig_oh, ig_ow; /* the geometry the tables were built for */
    int32_t *d_wzp; /* per-channel weight zero points when any is non-zero (asymmetric weights), else NULL */
    int8_t *d_lut; /* post table or the ACT table */
    int zp_in, zp_out, act, q6;
    float act_p0, act_p1; /* parameters of a unary op (leaky slope; clip min, max) */
    int binop;            /* B200_OPK_ADD: b200_binop (add / sub / mul) */
    int bcast_nc;         /* B200_OPK_ADD with a second ACTIVATION of shape [N, C, 1, 1]: one value per image and channel */
    void *d_const;        /* B200_OPK_ADD: constant second operand, one pixel's channels ([cp] elements); NULL = tensor */
    int const_count;      /* its length in elements */
    /* B200_OPK_CONCAT: device axis (0..3 = n, c, h, w), per-input offset along it and requant table */
    int cat_n, cat_axis;
    int cat_off[B200_CONCAT_MAX];
    int8_t *cat_lut[B200_CONCAT_MAX];
    /* the output qinfo the epilogue quantises to (needed when a relu is fused later) */
    float s_out;
    /* eltwise / pool / softmax */
    float s_in, s_in1;
    int zp_in1;
    int pool_avg, pool_global, count_include_pad;
    /* B200_OPK_TENSOR */
    int t_op;            /* enum b200_tensor_op */
    int t_axis, t_perm[4];
    float t_eps;
    int32_t *d_idx;      /* gather: constant indices */
    int n_idx, oob_q;
    float *d_gamma, *d_beta;
    int in_rank, in_dim[4], in1_rank, in1_dim[4], out_rank, out_dim[4]; /* logical shapes (the API's) */
    int mm_batches, mm_batches_b, mm_i, mm_k, mm_j, mm_trans_a, mm_trans_b, mm_const_b;
    int two_inputs;      /* the second operand is an activation (matmul of two activations) */
    /* layer-mode staging buffers, grown on demand */
    void *stg[8];
    size_t stg_bytes[8];
} b200_op;

enum b200_tensor_op { B200_T_TRANSPOSE = 1, B200_T_GATHER, B200_T_REDUCE_SUM, B200_T_LAYER_NORM, B200_T_RMS_NORM, B200_T_MATMUL };

/* registry params* -> op (ops.c) */
void b200_op_bind(void *params, b200_op *op);
b200_op *b200_op_find(void *params);

/* run on device tensors; scratch is im2col space (b200_op_scratch_bytes) */
const char *b200_op_kname(const b200_op *op, const b200_dt *in0); /* kernel that will actually run */
size_t b200_op_scratch_bytes(const b200_op *op, const b200_dt *in0, const b200_dt *out);
/* `part`: which input of a concat `in0` is (0 for every other kind) */
int b200_op_run(b200_op *op, int part, const b200_dt *in0, const b200_dt *in1, const b200_dt *out,
                void *scratch, void *stream);
/* depthwise 3x3 step + the 1x1 conv that is its only consumer as ONE kernel (csrc/dwpw_fused.cu);
 * `mid` carries the shape of the depthwise output, which is never materialised */
int b200_dwpw_can_fuse(const b200_op *dw, const b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out);
int b200_dwpw_prefers_fusion(b200_op *dw, b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out);
int b200_dwpw_run(b200_op *dw, b200_op *pw, const b200_dt *in, const b200_dt *mid, const b200_dt *out, void *stream);
/* fuse a following relu / relu6 node (with its own qinfo) into this op's epilogue */
int b200_op_can_fuse_act(const b200_op *op);
int b200_op_fuse_act(b200_op *op, int act, float p0, float p1, const struct csinn_tensor *act_in,
                     const struct csinn_tensor *act_out);

/* quant.c */
int b200_make_requant(b200_op *op, const struct csinn_tensor *input,
                      const struct csinn_tensor *kernel, const struct csinn_tensor *bias,
                      const struct csinn_tensor *output, int taps_per_o, int fuse_zp2bias,
                      int n_out);
struct csinn_tensor *b200_dequant_weights_f16(const struct csinn_tensor *kernel); /* fp16 activations, int8 weights */
struct csinn_tensor *b200_int8_weights_as_f16(const struct csinn_tensor *kernel); /* exact integers + scales for the epilogue */
void b200_free_dequant(struct csinn_tensor *t);
float b200_f16_to_f32(uint16_t h);
void *b200_pack_conv_weights(b200_op *op, const struct csinn_tensor *kernel, size_t *bytes);
void *b200_pack_dw_weights(b200_op *op, const struct csinn_tensor *kernel, int cp, size_t *bytes);
void *b200_pack_dw3x3_rows(b200_op *op, const struct csinn_tensor *kernel, int cp);
void *b200_pack_dw_diag(b200_op *op, const struct csinn_tensor *kernel, int cp, int *ldk);
int b200_make_igemm_tables(b200_op *op, const struct csinn_tensor *kernel, int h, int w, int oh, int ow);
void *b200_pack_fc_weights(b200_op *op, const struct csinn_tensor *weights, size_t *bytes);

/* graph.c */
typedef struct b200_graph b200_graph;
typedef struct b200_option {
    b200_ctx ctx;
    b200_graph *g;
} b200_option;
b200_option *b200_option_of(struct csinn_session *sess);

#endif
