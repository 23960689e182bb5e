// runtime.cu -- device / memory / stream / event / CUDA-graph plumbing of the b200nn C-ABI
// (include/b200nn.h).  No compute here.  Every entry point fails loudly: there is no CPU path.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include <stdlib.h>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_capturing_launches{0};
static thread_local bool g_capturing = false;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n)
{
    if (g_capturing)
        g_capturing_launches += n;
    else
        g_launches += n;
}

// Programmatic dependent launch.  Measured on B200 (MobileNetV1 int8, CUDA-graph replay): at batch 1 a
// step takes 138 us with it against 166 us without (every kernel's prologue -- barrier init, TMEM
// allocation, weight staging -- runs under its predecessor's tail), at batch 256 0.944 ms against
// 0.929 ms (early-resident CTAs only compete with the predecessor's last wave).  So the session
// captures its step list both ways and keeps the faster graph (b200_opt/graph.c); SHL_B200_PDL=0/1
// forces it.  Layer mode (one synchronous kernel per call) has nothing to overlap: off.
static int g_pdl = 0;
bool pdl_enabled() { return g_pdl != 0; }
extern "C" void b200_set_pdl(int on) { g_pdl = on ? 1 : 0; }
extern "C" int b200_get_pdl(void) { return g_pdl; }

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup: no link-time
// dependency on libcuda, so the library also loads (and exports its symbols) on a CPU box.
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn encode_entry()
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
            return nullptr;
        }
        fn = reinterpret_cast<encode_tiled_fn>(p);
    }
    return fn;
}

// pixel-major activation [n][h][w][cp] bytes as a 4-D tiled map (no swizzle, zero fill out of
// bounds -- negative start coordinates are how a conv halo is fetched)
int encode_tmap_nhwc_u8(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_c,
                        int box_w, int box_h)
{
    return encode_tmap_nhwc_u8_nb(map, base, n, h, w, cp, box_c, box_w, box_h, 1);
}

int encode_tmap_nhwc_u8_nb(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_c,
                           int box_w, int box_h, int box_n)
{
    encode_tiled_fn fn = encode_entry();
    if (!fn) return B200_ERR_CUDA;
    cuuint64_t gdim[4] = {(cuuint64_t)cp, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t gstride[3] = {(cuuint64_t)cp, (cuuint64_t)cp * w, (cuuint64_t)cp * w * h};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(4d) failed: CUresult %d (n %d h %d w %d cp %d box %d x %d x %d)", (int)r,
                  n, h, w, cp, box_c, box_w, box_h);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

// 4-D box of a pixel-major uint8 tensor with a choice of swizzle (0 / 32 / 64 / 128 bytes = the box's channel extent
// for the swizzled modes): halo loads and clipped output-tile stores of csrc/dwpw_fused.cu, csrc/conv_igemm.cu
int encode_tmap_nhwc_u8_ex(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_c, int box_w, int box_h,
                           int box_n, int swizzle_bytes)
{
    encode_tiled_fn fn = encode_entry();
    if (!fn) return B200_ERR_CUDA;
    cuuint64_t gdim[4] = {(cuuint64_t)cp, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t gstride[3] = {(cuuint64_t)cp, (cuuint64_t)cp * w, (cuuint64_t)cp * w * h};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                    : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                    : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                          : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(4d, swizzle %d) failed: CUresult %d (n %d h %d w %d cp %d box %d x %d x %d x %d)",
                  swizzle_bytes, (int)r, n, h, w, cp, box_c, box_w, box_h, box_n);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

// 4-D box {128 channels, box_w, box_h, box_n} of a pixel-major uint8 tensor, SWIZZLE_128B (csrc/dwconv3x3_umma128.cu)
int encode_tmap_nhwc_u8_sw128(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_w, int box_h,
                              int box_n)
{
    encode_tiled_fn fn = encode_entry();
    if (!fn) return B200_ERR_CUDA;
    cuuint64_t gdim[4] = {(cuuint64_t)cp, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t gstride[3] = {(cuuint64_t)cp, (cuuint64_t)cp * w, (cuuint64_t)cp * w * h};
    cuuint32_t box[4] = {128, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(4d, swizzle 128) failed: CUresult %d (n %d h %d w %d cp %d box %d x %d x %d)",
                  (int)r, n, h, w, cp, box_w, box_h, box_n);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

// im2col-mode tensor map over a pixel-major uint8 tensor [n][h][w][cp] (cuTensorMapEncodeIm2col): one load
// delivers `pixels` consecutive OUTPUT pixels of a convolution (walking W, then H, then N with the
// convolution's stride, the base pixel box shrunk by the corners so that exactly ow x oh base pixels exist per
// image) x `chans` channels of ONE filter tap, the tap given per load as (w, h) offsets; taps that fall outside
// the image are zero-filled.  lower = -pad, upper = pad_far - (k - 1) * dilation.
typedef CUresult (*encode_im2col_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                     const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);
int encode_tmap_im2col_u8(CUtensorMap *map, const void *base, int n, int h, int w, int c, int cp, int lower_w, int lower_h,
                          int upper_w, int upper_h, int stride_w, int stride_h, int chans, int pixels, int swizzle_bytes)
{
    static encode_im2col_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeIm2col entry point unavailable (%s)", cudaGetErrorString(e));
            return B200_ERR_CUDA;
        }
        fn = reinterpret_cast<encode_im2col_fn>(p);
    }
    cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t gstride[3] = {(cuuint64_t)cp, (cuuint64_t)cp * w, (cuuint64_t)cp * w * h};
    int lower[2] = {lower_w, lower_h}, upper[2] = {upper_w, upper_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(base), gdim, gstride, lower, upper,
                    (cuuint32_t)chans, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                    : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                    : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                          : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col failed: CUresult %d (n %d h %d w %d c %d cp %d corners %d %d %d %d stride %d %d "
                  "chans %d pixels %d)", (int)r, n, h, w, c, cp, lower_w, lower_h, upper_w, upper_h, stride_w, stride_h, chans,
                  pixels);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

int encode_tmap_2d(CUtensorMap *map, int elem_bytes, const void *base, uint64_t inner,
                   uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                   int swizzle_bytes)
{
    encode_tiled_fn fn = encode_entry();
    if (!fn) return B200_ERR_CUDA;
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstride[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt =
        elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(map, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                    : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                    : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                          : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: CUresult %d (inner %llu outer %llu pitch %llu box %u x %u)",
                  (int)r, (unsigned long long)inner, (unsigned long long)outer,
                  (unsigned long long)pitch_bytes, box_inner, box_outer);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_abi_version(void) { return B200NN_ABI_VERSION; }
const char *b200_last_error(void) { return g_err; }

int b200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
        return 0;
    }
    return n;
}

int b200_set_device(int dev)
{
    int n = b200_device_count();
    if (n <= 0) {
        set_error("no CUDA device visible: the b200 backend has no CPU fallback");
        return B200_ERR_NO_DEVICE;
    }
    B200_CUDA_CHECK(cudaSetDevice(dev));
    int major = 0;
    B200_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("device %d is sm_%d0, the kernels are built for sm_100a only", dev, major);
        return B200_ERR_NO_DEVICE;
    }
    return B200_OK;
}

int b200_sm_count(void) { return sm_count(); }

int b200_malloc(void **dptr, size_t bytes)
{
    if (!dptr) return B200_ERR_ARG;
    *dptr = nullptr;
    B200_CUDA_CHECK(cudaMalloc(dptr, bytes ? bytes : 16));
    return B200_OK;
}
int b200_free(void *dptr)
{
    if (dptr) B200_CUDA_CHECK(cudaFree(dptr));
    return B200_OK;
}
int b200_malloc_host(void **hptr, size_t bytes)
{
    if (!hptr) return B200_ERR_ARG;
    *hptr = nullptr;
    B200_CUDA_CHECK(cudaMallocHost(hptr, bytes ? bytes : 16));
    return B200_OK;
}
int b200_free_host(void *hptr)
{
    if (hptr) B200_CUDA_CHECK(cudaFreeHost(hptr));
    return B200_OK;
}
int b200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return B200_OK;
}
int b200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return B200_OK;
}
int b200_memcpy_d2h_rows(void *dst, const void *src, size_t row_bytes, size_t src_pitch, size_t rows, void *stream)
{
    B200_CUDA_CHECK(cudaMemcpy2DAsync(dst, row_bytes, src, src_pitch, row_bytes, rows, cudaMemcpyDeviceToHost,
                                      (cudaStream_t)stream));
    return B200_OK;
}
int b200_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
    B200_CUDA_CHECK(
        cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B200_OK;
}
int b200_memset(void *dst, int value, size_t bytes, void *stream)
{
    B200_CUDA_CHECK(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return B200_OK;
}
int b200_stream_create(void **stream)
{
    cudaStream_t s;
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return B200_OK;
}
int b200_stream_destroy(void *stream)
{
    if (stream) B200_CUDA_CHECK(cudaStreamDestroy((cudaStream_t)stream));
    return B200_OK;
}
int b200_stream_sync(void *stream)
{
    B200_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return B200_OK;
}
int b200_device_sync(void)
{
    B200_CUDA_CHECK(cudaDeviceSynchronize());
    return B200_OK;
}
int b200_event_create(void **ev)
{
    cudaEvent_t e;
    B200_CUDA_CHECK(cudaEventCreate(&e));
    *ev = e;
    return B200_OK;
}
int b200_event_destroy(void *ev)
{
    if (ev) B200_CUDA_CHECK(cudaEventDestroy((cudaEvent_t)ev));
    return B200_OK;
}
int b200_event_record(void *ev, void *stream)
{
    B200_CUDA_CHECK(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
    return B200_OK;
}
int b200_event_sync(void *ev)
{
    B200_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)ev));
    return B200_OK;
}
int b200_event_elapsed_ms(void *start, void *stop, float *ms)
{
    B200_CUDA_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return B200_OK;
}
int b200_stream_wait_event(void *stream, void *ev)
{
    B200_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
    return B200_OK;
}

// ---- CUDA graph capture -----------------------------------------------------------
struct GraphExec {
    cudaGraphExec_t exec;
    uint64_t launches;  // kernels recorded in the graph: counted on every replay
};

int b200_graph_begin(void *stream)
{
    g_capturing_launches = 0;
    B200_CUDA_CHECK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    g_capturing = true;
    return B200_OK;
}
int b200_graph_end(void *stream, void **graph_exec)
{
    cudaGraph_t g = nullptr;
    g_capturing = false;
    B200_CUDA_CHECK(cudaStreamEndCapture((cudaStream_t)stream, &g));
    GraphExec *ge = new GraphExec();
    ge->launches = g_capturing_launches.load();
    cudaError_t e = cudaGraphInstantiate(&ge->exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(e));
        delete ge;
        return B200_ERR_CUDA;
    }
    *graph_exec = ge;
    return B200_OK;
}
int b200_graph_launch(void *graph_exec, void *stream)
{
    GraphExec *ge = static_cast<GraphExec *>(graph_exec);
    B200_CUDA_CHECK(cudaGraphLaunch(ge->exec, (cudaStream_t)stream));
    g_launches += ge->launches;
    return B200_OK;
}
int b200_graph_destroy(void *graph_exec)
{
    GraphExec *ge = static_cast<GraphExec *>(graph_exec);
    if (ge) {
        cudaGraphExecDestroy(ge->exec);
        delete ge;
    }
    return B200_OK;
}

uint64_t b200_launch_count(void) { return g_launches.load(); }

int b200_flush_l2(void *stream)
{
    static void *scratch[64] = {nullptr};
    const size_t bytes = 256u << 20;  // 2 x the 126 MB L2
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return B200_ERR_ARG;
    if (!scratch[dev]) B200_CUDA_CHECK(cudaMalloc(&scratch[dev], bytes));
    static int v = 0;
    B200_CUDA_CHECK(cudaMemsetAsync(scratch[dev], ++v & 0xFF, bytes, (cudaStream_t)stream));
    return B200_OK;
}

// Host-side table for the `post` stage of the epilogue and for standalone relu / relu6
// nodes: the reference's float sequence (source/reference/utils.c:609, relu.c:20,
// relu6.c:20, source/nn2/utils.c:550) evaluated once per possible int8 input.
void b200_build_unary_lut(int8_t lut[256], int act, float p0, float p1, float s_in, int zp_in, float s_out, int zp_out)
{
    for (int q = -128; q < 128; q++) {
        volatile float d = (float)q - (float)zp_in;
        volatile float r = d * s_in;
        switch (act) {
            case B200_ACT_RELU:
                r = r > 0 ? r : 0;
                break;
            case B200_ACT_RELU6:
                r = r > 0 ? r : 0;
                r = (float)fmin(r, 6);
                break;
            case B200_ACT_LEAKY_RELU: /* val > 0 ? val : val * n */
                r = r > 0 ? r : r * p0;
                break;
            case B200_ACT_SIGMOID: /* 1.0f / (1.0f + exp(-val)): exp and the division in double, stored to float */
                r = (float)(1.0f / (1.0f + exp(-(double)r)));
                break;
            case B200_ACT_CLIP:
                r = r < p0 ? p0 : (r > p1 ? p1 : r);
                break;
            case B200_ACT_SILU: /* val / (1.0f + exp(-val)): double arithmetic, stored to float */
                r = (float)((double)r / (1.0f + exp(-(double)r)));
                break;
            case B200_ACT_ERF:
                r = (float)erf((double)r);
                break;
            default:
                break;
        }
        volatile float t = r / s_out;
        float v = (float)(nearbyint((double)t) + (double)zp_out);
        lut[q + 128] = v > 127 ? 127 : (v < -128 ? -128 : (int8_t)v);
    }
}

void b200_build_requant_lut(int8_t lut[256], int act, float s_in, int zp_in, float s_out,
                            int zp_out)
{
    b200_build_unary_lut(lut, act, 0.f, 0.f, s_in, zp_in, s_out, zp_out);
}

}  // extern "C"
