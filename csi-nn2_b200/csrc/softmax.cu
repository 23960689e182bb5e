// softmax.cu -- softmax over the channel axis of [rows][cp] (axis 1 of an NCHW tensor with
// H = W = 1), restating source/reference/softmax.c:20-66 step by step: f32 max, exp evaluated in
// double on the f32 difference, an f32 accumulator that receives the double terms ONE BY ONE in
// channel order (that order is what fixes the result bits, so one thread does it from values the
// whole block computed in parallel, through the integer restatement in softmax_sum.h), then
// double division narrowed to f32 and requantised.
#include <float.h>

#include "common.cuh"
#include "softmax_sum.h"

namespace b200 {

template <int DT>
__global__ void __launch_bounds__(256) softmax_kernel(const void *__restrict__ in_v,
                                                      void *__restrict__ out_v, int c, int cp_in,
                                                      int cp_out, float s_in, int zp_in,
                                                      float s_out, int zp_out)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    extern __shared__ double s_mem[];
    double *s_e = s_mem;                                   // [c]
    float *s_x = reinterpret_cast<float *>(s_e + c);       // [c]
    __shared__ float s_red[8];
    __shared__ float s_acc;
    const int row = blockIdx.x;
    const int tid = threadIdx.x;

    float mx = -FLT_MAX;
    for (int j = tid; j < c; j += blockDim.x) {
        float x;
        if (DT == B200_I8)
            x = dequant_i8(static_cast<const int8_t *>(in_v)[static_cast<long long>(row) * cp_in + j],
                           s_in, zp_in);
        else
            x = __half2float(static_cast<const __half *>(in_v)[static_cast<long long>(row) * cp_in + j]);
        s_x[j] = x;
        mx = fmaxf(mx, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int wv = 1; wv < 8; wv++) mx = fmaxf(mx, s_red[wv]);

    for (int j = tid; j < c; j += blockDim.x)
        s_e[j] = exp(static_cast<double>(__fsub_rn(s_x[j], mx)));
    __syncthreads();
    // the term-by-term float accumulation of the reference in its integer form (softmax_sum.h; same
    // bits): per binade of the running sum every thread turns its terms into integer increments,
    // then one thread adds them up -- a 4-cycle add per term where the literal float <- double step
    // costs ~50 (this chain is what the kernel's run time consists of)
    uint32_t *s_d = reinterpret_cast<uint32_t *>(s_x);  // the x values are dead from here on
    __shared__ int s_j;
    if (tid == 0) {
        s_acc = 0.f;
        s_j = 0;
    }
    __syncthreads();
    for (;;) {
        const float acc_now = s_acc;
        const int j0 = s_j;
        if (j0 >= c) break;
        const uint32_t ex = (__float_as_uint(acc_now) >> 23) & 0xFF;
        const bool normal = ex != 0 && ex < 0xFE;
        if (normal)
            for (int j = j0 + tid; j < c; j += blockDim.x) s_d[j] = b200_softmax_term(s_e[j], static_cast<int>(ex) - 127);
        __syncthreads();  // also: everybody has read s_acc / s_j
        if (tid == 0) {
            float a = acc_now;
            s_j = b200_softmax_chain(normal ? s_d : nullptr, s_e, j0, c, &a);
            s_acc = a;
        }
        __syncthreads();
    }
    const double acc = static_cast<double>(s_acc);
    for (int j = tid; j < cp_out; j += blockDim.x) {
        const float v = j < c ? static_cast<float>(s_e[j] / acc) : 0.f;
        if (DT == B200_I8)
            static_cast<int8_t *>(out_v)[static_cast<long long>(row) * cp_out + j] =
                j < c ? static_cast<int8_t>(quant_i8_exact(v, s_out, zp_out)) : 0;
        else
            static_cast<__half *>(out_v)[static_cast<long long>(row) * cp_out + j] =
                __float2half_rn(v);
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_softmax(int dtype, const void *in, void *out, int rows, int c, int cp_in,
                            int cp_out, float s_in, int zp_in, float s_out, int zp_out,
                            void *stream)
{
    if ((dtype != B200_I8 && dtype != B200_F16) || !in || !out || rows <= 0 || c <= 0 ||
        cp_in < c || cp_out < c) {
        set_error("b200_softmax: bad arguments (dtype=%d rows=%d c=%d)", dtype, rows, c);
        return B200_ERR_ARG;
    }
    const size_t smem = static_cast<size_t>(c) * (sizeof(double) + sizeof(float));
    if (smem > 48 * 1024) {
        set_error("b200_softmax: axis length %d exceeds the 4096-channel kernel limit", c);
        return B200_ERR_UNSUPPORTED;
    }
    if (dtype == B200_I8)
        launch_kernel(softmax_kernel<B200_I8>, dim3(rows), dim3(256), smem, (cudaStream_t)stream, 
            in, out, c, cp_in, cp_out, s_in, zp_in, s_out, zp_out);
    else
        launch_kernel(softmax_kernel<B200_F16>, dim3(rows), dim3(256), smem, (cudaStream_t)stream, 
            in, out, c, cp_in, cp_out, s_in, zp_in, s_out, zp_out);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
