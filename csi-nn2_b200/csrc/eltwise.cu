// eltwise.cu -- bandwidth-bound elementwise ops on 128-bit vectors.
//   b200_lut_i8   : relu / relu6 / requantising identity as a 256-entry table (the reference's
//                   dequant -> f32 op -> requant is a pure function of the int8 input:
//                   source/reference/relu.c:39, relu6.c:42, utils.c:609)
//   b200_relu_f16 : fp16 relu / relu6
//   b200_add      : elementwise add with per-tensor qinfo (source/reference/add.c:36 through
//                   diso_callback_base, utils.c:622), exact float sequence incl. IEEE division
#include "common.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) lut_i8_kernel(const uint4 *__restrict__ in,
                                                     uint4 *__restrict__ out, long long nvec,
                                                     const int8_t *__restrict__ lut)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = static_cast<uint8_t>(lut[i]);
    __syncthreads();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 v = __ldg(in + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t o = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t idx = ((w[q] >> (8 * e)) & 0xFF) ^ 0x80;  // q + 128
                o |= static_cast<uint32_t>(s_lut[idx]) << (8 * e);
            }
            r[q] = o;
        }
        out[i] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

__global__ void __launch_bounds__(256) relu_f16_kernel(const uint4 *__restrict__ in,
                                                       uint4 *__restrict__ out, long long nvec,
                                                       int act)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    const __half2 zero = __float2half2_rn(0.f), six = __float2half2_rn(6.f);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        uint4 v = __ldg(in + i);
        __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (act != B200_ACT_NONE) h[q] = __hmax2(h[q], zero);
            if (act == B200_ACT_RELU6) h[q] = __hmin2(h[q], six);
        }
        out[i] = v;
    }
}

// the unary ops without a packed-half form: f32 arithmetic on the converted value (what the
// reference's fp16 path does: convert, f32 op, convert back -- utils.c:609)
__global__ void __launch_bounds__(256) unary_f16_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                        long long nvec, int act, float p0, float p1)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        uint4 v = __ldg(in + i);
        __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float2 f = __half22float2(h[q]);
            float *x = &f.x;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                float r = x[e];
                if (act == B200_ACT_LEAKY_RELU) r = r > 0.f ? r : r * p0;
                else if (act == B200_ACT_SIGMOID) r = static_cast<float>(1.0 / (1.0 + exp(-static_cast<double>(r))));
                else if (act == B200_ACT_CLIP) r = r < p0 ? p0 : (r > p1 ? p1 : r);
                else r = act_f(r, act);
                x[e] = r;
            }
            h[q] = __floats2half2_rn(f.x, f.y);
        }
        out[i] = v;
    }
}

struct AddArgs {
    float s_a, s_b, s_out;
    int zp_a, zp_b, zp_out, act;
    const int8_t *post_lut;
    int binop;  // b200_binop: add / sub / mul (source/reference/add.c, sub.c, mul.c: one f32 op between the dequantised values)
};
__device__ __forceinline__ float binop_f(float a, float b, int binop)
{
    return binop == B200_BINOP_SUB ? __fsub_rn(a, b) : (binop == B200_BINOP_MUL ? __fmul_rn(a, b) : __fadd_rn(a, b));
}

__global__ void __launch_bounds__(256) add_i8_kernel(const uint4 *__restrict__ a,
                                                     const uint4 *__restrict__ b,
                                                     uint4 *__restrict__ out, long long nvec,
                                                     const AddArgs p)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __shared__ uint8_t s_lut[256];
    if (p.post_lut != nullptr)
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
            s_lut[i] = static_cast<uint8_t>(p.post_lut[i]);
    __syncthreads();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 va = __ldg(a + i), vb = __ldg(b + i);
        const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t o = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int qa = static_cast<int8_t>(wa[q] >> (8 * e));
                const int qb = static_cast<int8_t>(wb[q] >> (8 * e));
                const float f = binop_f(dequant_i8(qa, p.s_a, p.zp_a), dequant_i8(qb, p.s_b, p.zp_b), p.binop);
                int qo = quant_i8_exact(f, p.s_out, p.zp_out);
                if (p.post_lut != nullptr) qo = static_cast<int8_t>(s_lut[qo + 128]);
                o |= (static_cast<uint32_t>(qo) & 0xFF) << (8 * e);
            }
            r[q] = o;
        }
        out[i] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

__global__ void __launch_bounds__(256) add_f16_kernel(const uint4 *__restrict__ a,
                                                      const uint4 *__restrict__ b,
                                                      uint4 *__restrict__ out, long long nvec,
                                                      int act, int binop)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 va = __ldg(a + i), vb = __ldg(b + i);
        const __half2 *ha = reinterpret_cast<const __half2 *>(&va);
        const __half2 *hb = reinterpret_cast<const __half2 *>(&vb);
        uint4 vo;
        __half2 *ho = reinterpret_cast<__half2 *>(&vo);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float2 fa = __half22float2(ha[q]), fb = __half22float2(hb[q]);
            ho[q] = __floats2half2_rn(act_f(binop_f(fa.x, fb.x, binop), act), act_f(binop_f(fa.y, fb.y, binop), act));
        }
        out[i] = vo;
    }
}

static int ew_grid(long long nvec)
{
    long long g = (nvec + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b200

using namespace b200;

extern "C" int b200_lut_i8(const void *in, void *out, size_t count, const int8_t *lut_dev,
                           void *stream)
{
    if (!in || !out || !lut_dev || count == 0 || count % 16 || !aligned16(in) || !aligned16(out)) {
        set_error("b200_lut_i8: bad arguments (count=%zu must be a non-zero multiple of 16)", count);
        return B200_ERR_ARG;
    }
    const long long nvec = static_cast<long long>(count / 16);
    launch_kernel(lut_i8_kernel, dim3(ew_grid(nvec)), dim3(256), 0, (cudaStream_t)stream, 
        static_cast<const uint4 *>(in), static_cast<uint4 *>(out), nvec, lut_dev);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_relu_f16(const void *in, void *out, size_t count, int act, void *stream)
{
    if (!in || !out || count == 0 || count % 8 || !aligned16(in) || !aligned16(out)) {
        set_error("b200_relu_f16: bad arguments (count=%zu must be a non-zero multiple of 8)", count);
        return B200_ERR_ARG;
    }
    const long long nvec = static_cast<long long>(count / 8);
    launch_kernel(relu_f16_kernel, dim3(ew_grid(nvec)), dim3(256), 0, (cudaStream_t)stream, 
        static_cast<const uint4 *>(in), static_cast<uint4 *>(out), nvec, act);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_unary_f16(const void *in, void *out, size_t count, int act, float p0, float p1, void *stream)
{
    if (!in || !out || count == 0 || count % 8 || !aligned16(in) || !aligned16(out)) {
        set_error("b200_unary_f16: bad arguments (count=%zu must be a non-zero multiple of 8)", count);
        return B200_ERR_ARG;
    }
    const long long nvec = static_cast<long long>(count / 8);
    launch_kernel(unary_f16_kernel, dim3(ew_grid(nvec)), dim3(256), 0, (cudaStream_t)stream,
                  static_cast<const uint4 *>(in), static_cast<uint4 *>(out), nvec, act, p0, p1);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_add(int dtype, const void *a, const void *b, void *out, size_t count, float s_a,
                        int zp_a, float s_b, int zp_b, float s_out, int zp_out,
                        const int8_t *post_lut, int act, void *stream)
{
    return b200_binary(B200_BINOP_ADD, dtype, a, b, out, count, s_a, zp_a, s_b, zp_b, s_out, zp_out, post_lut, act, stream);
}

extern "C" int b200_binary(int binop, int dtype, const void *a, const void *b, void *out, size_t count, float s_a,
                           int zp_a, float s_b, int zp_b, float s_out, int zp_out,
                           const int8_t *post_lut, int act, void *stream)
{
    if (binop < B200_BINOP_ADD || binop > B200_BINOP_MUL) {
        set_error("b200_binary: unknown op %d", binop);
        return B200_ERR_ARG;
    }
    const int vec = dtype == B200_I8 ? 16 : 8;
    if ((dtype != B200_I8 && dtype != B200_F16) || !a || !b || !out || count == 0 || count % vec ||
        !aligned16(a) || !aligned16(b) || !aligned16(out)) {
        set_error("b200_add: bad arguments (dtype=%d count=%zu)", dtype, count);
        return B200_ERR_ARG;
    }
    const long long nvec = static_cast<long long>(count / vec);
    if (dtype == B200_I8) {
        AddArgs p{s_a, s_b, s_out, zp_a, zp_b, zp_out, act, post_lut, binop};
        launch_kernel(add_i8_kernel, dim3(ew_grid(nvec)), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const uint4 *>(a), static_cast<const uint4 *>(b), static_cast<uint4 *>(out),
            nvec, p);
    } else {
        launch_kernel(add_f16_kernel, dim3(ew_grid(nvec)), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const uint4 *>(a), static_cast<const uint4 *>(b), static_cast<uint4 *>(out),
            nvec, act, binop);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}
