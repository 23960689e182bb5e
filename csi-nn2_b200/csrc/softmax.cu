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

// Denominator of one row: sum of the c non-negative doubles in s_e, accumulated exactly like the
// reference's `float acc += double` loop (same bits).  Called by all threads of the block; returns the
// sum to every thread.
//
// The reference's loop is a 1000-step dependent chain of float <- double conversions (~27 us, most
// of this kernel and a fifth of a batch-1 inference).  In its integer form (softmax_sum.h) every
// term is, inside one binade of the running sum, an integer increment of the float mantissa, and
// integer adds commute: warp 0 takes windows of 64 terms (two per lane), prefix-scans their
// increments, and either consumes the whole window or stops at the first term at which the sum
// would leave the binade (or that is a rounding tie) -- that one term takes the literal step and the
// scan resumes behind it in the new binade.  No block barrier inside the loop.
__device__ float block_softmax_denominator(const double *s_e, int c)
{
    __shared__ float s_acc;
    const int tid = threadIdx.x;
    if (tid < 32) {
        const int lane = tid;
        float acc = 0.f;
        int j = 0;
        while (j < c) {
            const uint32_t abits = __float_as_uint(acc);
            const uint32_t ex = (abits >> 23) & 0xFF;
            if (ex == 0 || ex >= 0xFE) {  // zero / subnormal / overflowing sum: one literal step (all lanes alike)
                acc = b200_softmax_sum_literal(s_e + j, 1, acc);
                j++;
                continue;
            }
            const uint32_t a_mant = (abits & 0x7FFFFFu) | 0x800000u;
            const unsigned long long room = (1ull << 24) - a_mant;  // the increments consumed must stay below this
            const int i0 = j + 2 * lane;
            const uint32_t v0 = i0 < c ? b200_softmax_term(s_e[i0], static_cast<int>(ex) - 127) : 0u;
            const uint32_t v1 = i0 + 1 < c ? b200_softmax_term(s_e[i0 + 1], static_cast<int>(ex) - 127) : 0u;
            // a term that must be literal always "leaves the binade"
            const unsigned long long w0 = (v0 & B200_SOFTMAX_LITERAL) ? (1ull << 40) : v0;
            const unsigned long long w1 = (v1 & B200_SOFTMAX_LITERAL) ? (1ull << 40) : v1;
            unsigned long long incl = w0 + w1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const unsigned long long before = incl - (w0 + w1);
            const bool stop0 = before + w0 >= room, stop1 = before + w0 + w1 >= room;
            const uint32_t stops = __ballot_sync(0xffffffffu, stop0 || stop1);
            if (stops == 0) {
                const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
                acc = __uint_as_float((ex << 23) | ((a_mant + static_cast<uint32_t>(total)) & 0x7FFFFFu));
                j += 64;
                continue;
            }
            const int who = __ffs(stops) - 1;
            const int p = __shfl_sync(0xffffffffu, stop0 ? i0 : i0 + 1, who);
            const unsigned long long used = __shfl_sync(0xffffffffu, stop0 ? before : before + w0, who);
            acc = __uint_as_float((ex << 23) | ((a_mant + static_cast<uint32_t>(used)) & 0x7FFFFFu));
            acc = b200_softmax_sum_literal(s_e + p, 1, acc);
            j = p + 1;
        }
        if (lane == 0) s_acc = acc;
    }
    __syncthreads();
    return s_acc;
}

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
    const float denom = block_softmax_denominator(s_e, c);
    const double acc = static_cast<double>(denom);
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

// test hook: the denominator code alone, one row per block
__global__ void __launch_bounds__(256) softmax_denominator_test_kernel(const double *e, int c, float *out)
{
    extern __shared__ double s_mem[];
    double *s_e = s_mem;
    for (int j = threadIdx.x; j < c; j += blockDim.x) s_e[j] = e[static_cast<size_t>(blockIdx.x) * c + j];
    __syncthreads();
    const float d = block_softmax_denominator(s_e, c);
    if (threadIdx.x == 0) out[blockIdx.x] = d;
}

}  // namespace b200

using namespace b200;

// TEST HOOK (tests/test_gpu_parity.py): rows x c non-negative doubles on the device -> rows float sums
// through the softmax kernel's denominator code; compared with the literal loop on the host
extern "C" int b200_test_softmax_denominator(const void *e_dev, int rows, int c, void *out_dev, void *stream)
{
    const size_t smem = static_cast<size_t>(c) * (sizeof(double) + sizeof(float));
    if (!e_dev || !out_dev || rows <= 0 || c <= 0 || smem > 48 * 1024) {
        set_error("b200_test_softmax_denominator: bad arguments");
        return B200_ERR_ARG;
    }
    launch_kernel(softmax_denominator_test_kernel, dim3(rows), dim3(256), smem, (cudaStream_t)stream,
                  static_cast<const double *>(e_dev), c, static_cast<float *>(out_dev));
    B200_LAUNCH_CHECK();
    return B200_OK;
}

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
