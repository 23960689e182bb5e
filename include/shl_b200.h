/*
 * shl_b200.h -- host-side entry points of the B200 backend for the CSI-NN2 / SHL C API.
 *
 * This is the reference-facing half of the drop-in boundary: C, compiled together with the
 * reference's own unchanged dispatch layer (source/nn2, source/graph_ref, source/utils) into
 * libshl_b200.so.  A user of the reference keeps calling csinn_conv2d_init / csinn_conv2d /
 * csinn_session_setup / csinn_session_run ...; the functions below are what the registry
 * (source/nn2/setup.c:98-129) hands those calls to.  They mirror, name for name, what a
 * reference backend exports (compare source/c920_opt/setup.c:24-386 and
 * include/backend/rvv/rvv.h); the device work is done through include/b200nn.h.
 *
 * Registration.  The reference's shl_init() (source/nn2/setup.c:36-71) calls
 * shl_target_init_rvv / _c906 / _c908 / _c920 / _c920v2 when built with the matching
 * SHL_BUILD_* macro.  libshl_b200.so defines those five symbols (the RISC-V directories are not
 * compiled) and each registers the b200 op map + runtime map under its own api id
 * (CSINN_RVV=15, CSINN_C906=3, CSINN_C908=12, CSINN_C920=4, CSINN_C920V2=18;
 * include/csinn/csinn_data_structure.h:94-115), so e.g. example/c906_mobilenetv1_f16.c
 * (sess->base_api = CSINN_C906) runs on the GPU unmodified.  There is no fall-through to
 * shl_cb_map_ref: an (op, dtype) pair b200 does not implement maps to an all-NULL callback and
 * csinn_<op>() returns CSINN_CALLBACK_UNSET (source/nn2/convolution.c:81).
 */
#ifndef SHL_B200_H_
#define SHL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct csinn_tensor;
struct csinn_session;
struct csinn_callback;
struct csinn_conv2d_params;
struct csinn_fc_params;
struct csinn_relu_params;
struct csinn_diso_params;
struct csinn_pool_params;
struct csinn_softmax_params;
struct csinn_reshape_params;
struct csinn_flatten_params;
struct csinn_perf_info;

/* ---- registration (replaces source/thead_rvv/setup.c:68 and source/c9*_opt/setup.c) ------ */
void shl_target_init_b200(int api); /* register the maps under one api id */
void shl_target_init_rvv(void);
void shl_target_init_c906(void);
void shl_target_init_c908(void);
void shl_target_init_c920(void);
void shl_target_init_c920v2(void);
struct csinn_callback *shl_cb_map_b200(int op, int dtype); /* cf. shl_cb_map_rvv, thead_rvv/setup.c:43 */
void *shl_b200_runtime_callback(int api);                  /* cf. shl_c920_runtime_callback, c920_opt/setup.c:354 */

/* ---- operator callbacks: init / exec pairs (cf. include/backend/rvv/rvv.h) --------------- */
int shl_b200_conv2d_init(struct csinn_tensor *input, struct csinn_tensor *output,
                         struct csinn_tensor *kernel, struct csinn_tensor *bias,
                         struct csinn_conv2d_params *params);
int shl_b200_conv2d(struct csinn_tensor *input, struct csinn_tensor *output,
                    struct csinn_tensor *kernel, struct csinn_tensor *bias,
                    struct csinn_conv2d_params *params);
int shl_b200_depthwise_conv2d_init(struct csinn_tensor *input, struct csinn_tensor *output,
                                   struct csinn_tensor *kernel, struct csinn_tensor *bias,
                                   struct csinn_conv2d_params *params);
int shl_b200_depthwise_conv2d(struct csinn_tensor *input, struct csinn_tensor *output,
                              struct csinn_tensor *kernel, struct csinn_tensor *bias,
                              struct csinn_conv2d_params *params);
int shl_b200_fullyconnected_init(struct csinn_tensor *input, struct csinn_tensor *output,
                                 struct csinn_tensor *weights, struct csinn_tensor *bias,
                                 struct csinn_fc_params *params);
int shl_b200_fullyconnected(struct csinn_tensor *input, struct csinn_tensor *output,
                            struct csinn_tensor *weights, struct csinn_tensor *bias,
                            struct csinn_fc_params *params);
int shl_b200_relu_init(struct csinn_tensor *input, struct csinn_tensor *output,
                       struct csinn_relu_params *params);
/* leaky relu / sigmoid / clip share shl_b200_relu as exec; their init callbacks */
void *shl_b200_sub_init_fn(void); /* sub / mul share shl_b200_add as exec */
void *shl_b200_mul_init_fn(void);
void *shl_b200_leaky_relu_init_fn(void);
void *shl_b200_sigmoid_init_fn(void);
void *shl_b200_clip_init_fn(void);
void *shl_b200_global_maxpool_init_fn(void); /* csinn_global_maxpool2d, source/reference/global_maxpool.c:21 */
void *shl_b200_div_init_fn(void);   /* csinn_div, source/reference/div.c:36 */
/* transpose / gather / reduce_sum / layer_norm / rms_norm / matmul (source/thead_rvv/setup.c:316-470; semantics
 * source/reference/transpose.c, gather.c, reduce_sum.c, layer_norm.c, rms_norm.c, matmul.c) */
void *shl_b200_transpose_init_fn(void);
void *shl_b200_gather_init_fn(void);
void *shl_b200_reduce_sum_init_fn(void);
void *shl_b200_layer_norm_init_fn(void);
void *shl_b200_rms_norm_init_fn(void);
void *shl_b200_matmul_init_fn(void);
void *shl_b200_tensor_exec1_fn(void);
void *shl_b200_gather_exec_fn(void);
void *shl_b200_norm_exec4_fn(void);
void *shl_b200_rms_norm_exec_fn(void);
void *shl_b200_matmul_exec_fn(void);
void *shl_b200_prelu_init_fn(void); /* csinn_prelu with a constant per-channel slope, source/reference/prelu.c:21 */
void *shl_b200_silu_init_fn(void); /* csinn_silu: val / (1 + exp(-val)), source/reference/silu.c:21 */
void *shl_b200_erf_init_fn(void);  /* csinn_erf, source/reference/erf.c:21 */
int shl_b200_relu(struct csinn_tensor *input, struct csinn_tensor *output,
                  struct csinn_relu_params *params);
int shl_b200_add_init(struct csinn_tensor *input0, struct csinn_tensor *input1,
                      struct csinn_tensor *output, struct csinn_diso_params *params);
int shl_b200_add(struct csinn_tensor *input0, struct csinn_tensor *input1,
                 struct csinn_tensor *output, struct csinn_diso_params *params);
int shl_b200_pool2d_init(struct csinn_tensor *input, struct csinn_tensor *output,
                         struct csinn_pool_params *params);
int shl_b200_pool2d(struct csinn_tensor *input, struct csinn_tensor *output,
                    struct csinn_pool_params *params);
int shl_b200_softmax_init(struct csinn_tensor *input, struct csinn_tensor *output,
                          struct csinn_softmax_params *params);
int shl_b200_softmax(struct csinn_tensor *input, struct csinn_tensor *output,
                     struct csinn_softmax_params *params);
/* concat along any axis of rank 1..4 tensors (replaces shl_rvv_concat_int8 / shl_rvv_concat_fp16,
 * source/thead_rvv/setup.c; semantics source/reference/concat.c:52) */
int shl_b200_concat_init(struct csinn_tensor **input, struct csinn_tensor *output,
                         struct csinn_concat_params *params);
int shl_b200_concat(struct csinn_tensor **input, struct csinn_tensor *output, struct csinn_concat_params *params);
/* split along any axis of rank 1..4 tensors (source/reference/split.c:81; shl_gref_split in the RVV table) */
int shl_b200_split_init(struct csinn_tensor *input, struct csinn_tensor **output, struct csinn_split_params *params);
int shl_b200_split(struct csinn_tensor *input, struct csinn_tensor **output, struct csinn_split_params *params);
int shl_b200_reshape_init(struct csinn_tensor *input, struct csinn_tensor *output, void *params);
int shl_b200_reshape(struct csinn_tensor *input, struct csinn_tensor *output, void *params);
/* perf callbacks: kernel name for the trace profiler (cf. source/thead_rvv/performance.c:442);
 * one per argument-list shape of source/graph_ref/setup.c:509-540 */
int shl_b200_perf(struct csinn_tensor *input, struct csinn_tensor *output,
                  struct csinn_tensor *kernel, struct csinn_tensor *bias, void *params,
                  struct csinn_perf_info *perf_info);
int shl_b200_perf_siso(struct csinn_tensor *input, struct csinn_tensor *output, void *params,
                       struct csinn_perf_info *perf_info);
int shl_b200_perf_diso(struct csinn_tensor *input0, struct csinn_tensor *input1,
                       struct csinn_tensor *output, void *params,
                       struct csinn_perf_info *perf_info);

/* ---- session / graph runtime (cf. shl_c920_session_*, shl_gref_session_run) -------------- */
void shl_b200_session_init(struct csinn_session *sess);
void shl_b200_session_deinit(struct csinn_session *sess);
int shl_b200_session_setup(struct csinn_session *sess);
int shl_b200_session_run(struct csinn_session *sess);
/* CSINN_LOAD_BG: the HHB binary model format (csinn_import_binary_model / csinn_load_binary_model,
 * source/nn2/format.c:1304, source/nn2/setup.c:546; cf. shl_gref_load_binary_model,
 * source/graph_ref/setup.c:929).  session_setup writes the same format when
 * sess->model.save_mode asks for it (source/graph_ref/setup.c:733-855). */
int shl_b200_load_binary_model(struct csinn_session *sess);

/* ---- b200-specific session controls (additions; everything above is the reference's API) - */
/* device ordinal for sessions created afterwards (default: $LOCAL_RANK, else 0) */
int shl_b200_set_device(int device);
/* Replay the session's captured CUDA graph on inputs already resident in HBM: no H2D, no D2H,
 * asynchronous.  `csinn_session_run` = update H2D + this + D2H + sync. */
int shl_b200_session_launch(struct csinn_session *sess);
int shl_b200_session_sync(struct csinn_session *sess);
void *shl_b200_session_stream(struct csinn_session *sess);
/* number of device kernels one session_run launches, and the fused step list for inspection */
int shl_b200_session_num_kernels(struct csinn_session *sess);
int shl_b200_session_describe(struct csinn_session *sess, char *buf, int buflen);
/* Input prefetch: start the H2D of the NEXT batch (host_ptr, pinned) on a copy stream so that it
 * overlaps the current csinn_session_run; the following csinn_update_input with the same pointer
 * then costs a device-side copy instead of an H2D.  Every batch still crosses PCIe once. */
int shl_b200_session_prefetch_input(int index, const void *host_ptr, struct csinn_session *sess);
int shl_b200_update_input(int index, struct csinn_tensor *input, struct csinn_session *sess);
/* per-step device time (ms, CUDA events on the session stream) with the algorithmic bytes / ops of
 * each step; the device-side counterpart of shl_benchmark_layer (source/utils/debug.c:1037).
 * Returns the number of steps written (<= cap). */
int shl_b200_session_profile(struct csinn_session *sess, int warmup, int iters, double *ms, double *bytes,
                             double *ops, int cap);
/* Weight arena of a set-up session (packed weights + per-channel tables, one contiguous
 * device allocation): exposed so that multi-GPU launchers can broadcast it once over NCCL
 * instead of re-uploading per rank. */
int shl_b200_session_weight_arena(struct csinn_session *sess, void **dev_ptr, uint64_t *bytes);
/* Multi-GPU start-up in C (b200_opt/dist.c): batch sharding has no collective on the inference path; the ONE
 * exchange is the weight arena of rank `root` into the arenas of the other ranks (created with
 * SHL_B200_SKIP_WEIGHT_UPLOAD=1) by a single ncclBroadcast over NVLink.  NCCL is bound at run time (dlopen), the
 * communicator is the library's own: shl_b200_nccl_unique_id on one rank, the 128 bytes carried to the others by
 * the host's launcher (MPI, torch.distributed, a file), shl_b200_nccl_comm_init everywhere. */
int shl_b200_nccl_unique_id(void *id128);
int shl_b200_nccl_comm_init(struct csinn_session *sess, const void *id128, int rank, int world, void **comm);
int shl_b200_session_broadcast_weights(struct csinn_session *sess, void *comm, int root);
int shl_b200_nccl_comm_destroy(void *comm);
/* drop the device operator bound to a params struct (staging buffers of layer mode) */
void shl_b200_op_release(void *params);
/* Errors.  The reference's front ends drop the status init / exec return
 * (source/nn2/convolution.c:50-55,64-86), so failures are also printed to stderr, kept here and
 * counted; SHL_B200_ABORT_ON_ERROR=1 in the environment turns them into abort(). */
const char *shl_b200_last_error(void);
int shl_b200_error_count(void);
void shl_b200_clear_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SHL_B200_H_ */
