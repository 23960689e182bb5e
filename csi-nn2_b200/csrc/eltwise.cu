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
                else if (act == B200_ACT_SILU) r = static_cast<float>(static_cast<double>(r) / (1.0 + exp(-static_cast<double>(r))));
                else if (act == B200_ACT_ERF) r = static_cast<float>(erf(static_cast<double>(r)));
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
    float inv_out;  // RN(1 / s_out): decides rint() of the quotient except near half-integers
};
__device__ __forceinline__ float binop_f(float a, float b, int binop)
{
    if (binop == B200_BINOP_PRELU) return a >= 0.f ? a : __fmul_rn(a, b);
    if (binop == B200_BINOP_DIV) return __fdiv_rn(a, b);
    return binop == B200_BINOP_SUB ? __fsub_rn(a, b) : (binop == B200_BINOP_MUL ? __fmul_rn(a, b) : __fadd_rn(a, b));
}

// int8 binary op, same bits as the reference's float sequence (dequantise both, one f32 op, IEEE
// division by s_out, round half even, + zp_out, clamp) at a third of its instruction count:
//  * byte -> float by PRMT into the mantissa of 1.5 * 2^23 and one packed FADD (exact integers);
//  * packed f32x2 arithmetic;
//  * the division is needed only to decide rint(): t = r * RN(1 / s_out) is within 2^-14 of the
//    IEEE quotient wherever the result does not saturate, so rint(t) is rint(quotient) unless t sits
//    within 2^-11 of a half-integer -- those elements (about 0.1 %) take the real __fdiv_rn.
__device__ __forceinline__ uint64_t bytes_to_f2(uint32_t wx, int e0, float off)
{
    // wx = word ^ 0x80808080: byte e = q + 128; bits 0x4B4000uu = 12582912 + (q + 128) as a float
    const uint32_t lo = __byte_perm(wx, 0x4B400000u, 0x7650 + e0);        // byte e0     -> low byte
    const uint32_t hi = __byte_perm(wx, 0x4B400000u, 0x7650 + e0 + 1);    // byte e0 + 1 -> low byte
    return f2_add(f2_pack_bits(lo, hi), f2_pack(off, off));               // q - zp, exact
}

__global__ void __launch_bounds__(256) add_i8_kernel(const uint4 *__restrict__ a,
                                                     const uint4 *__restrict__ b,
                                                     uint4 *__restrict__ out, long long nvec,
                                                     const AddArgs p, const int b_period, const long long b_outer)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __shared__ uint8_t s_lut[256];
    const bool has_lut = p.post_lut != nullptr;
    if (has_lut)
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
            s_lut[i] = static_cast<uint8_t>(p.post_lut[i]);
    __syncthreads();
    const float off_a = -(kMagicF + 128.f + static_cast<float>(p.zp_a)), off_b = -(kMagicF + 128.f + static_cast<float>(p.zp_b));
    const uint64_t sa2 = f2_pack(p.s_a, p.s_a), sb2 = f2_pack(p.s_b, p.s_b), inv2 = f2_pack(p.inv_out, p.inv_out);
    const uint64_t mg2 = f2_pack(kMagicF, kMagicF), nmg2 = f2_pack(-kMagicF, -kMagicF);
    const float zo = static_cast<float>(p.zp_out);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 va = __ldg(a + i), vb = __ldg(b + (b_period ? (b_outer ? i / b_outer * b_period : 0) + i % b_period : i));
        const uint32_t wa[4] = {va.x ^ 0x80808080u, va.y ^ 0x80808080u, va.z ^ 0x80808080u, va.w ^ 0x80808080u};
        const uint32_t wb[4] = {vb.x ^ 0x80808080u, vb.y ^ 0x80808080u, vb.z ^ 0x80808080u, vb.w ^ 0x80808080u};
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int qo[4];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                // dequantise two elements of each operand: (q - zp) * s, the reference's two roundings
                const uint64_t xa = f2_fma(bytes_to_f2(wa[q], 2 * h, off_a), sa2, 0ull);
                const uint64_t xb = f2_fma(bytes_to_f2(wb[q], 2 * h, off_b), sb2, 0ull);
                uint64_t rr;
                if (p.binop == B200_BINOP_DIV) {
                    int a0, a1, d0, d1;
                    f2_unpack_bits(xa, a0, a1);
                    f2_unpack_bits(xb, d0, d1);
                    rr = f2_pack(__fdiv_rn(__int_as_float(a0), __int_as_float(d0)), __fdiv_rn(__int_as_float(a1), __int_as_float(d1)));
                } else if (p.binop == B200_BINOP_PRELU) {
                    int a0, a1, m0, m1;
                    f2_unpack_bits(xa, a0, a1);
                    f2_unpack_bits(f2_fma(xa, xb, 0ull), m0, m1);
                    // input >= 0 keeps the input (-0.0 included: it compares equal to 0)
                    rr = f2_pack_bits(static_cast<uint32_t>(__int_as_float(a0) >= 0.f ? a0 : m0),
                                      static_cast<uint32_t>(__int_as_float(a1) >= 0.f ? a1 : m1));
                } else if (p.binop == B200_BINOP_MUL)
                    rr = f2_fma(xa, xb, 0ull);
                else if (p.binop == B200_BINOP_SUB)
                    rr = f2_add(xa, f2_pack_bits(static_cast<uint32_t>(xb) ^ 0x80000000u, static_cast<uint32_t>(xb >> 32) ^ 0x80000000u));
                else
                    rr = f2_add(xa, xb);
                const uint64_t t2 = f2_fma(rr, inv2, 0ull);
                const uint64_t n2 = f2_add(f2_add(t2, mg2), nmg2);  // round half even (|t| < 2^22; beyond it saturates anyway)
                int rb0, rb1, tb0, tb1, nb0, nb1;
                f2_unpack_bits(rr, rb0, rb1);
                f2_unpack_bits(t2, tb0, tb1);
                f2_unpack_bits(n2, nb0, nb1);
                float n0 = __int_as_float(nb0), n1 = __int_as_float(nb1);
                const float t0 = __int_as_float(tb0), t1 = __int_as_float(tb1);
                // within 2^-11 of a half-integer (or too large for the magic rounding): the real division decides
                if (fabsf(t0 - n0) > 0.49951171875f || !(fabsf(t0) < 4194304.f)) n0 = rintf(__fdiv_rn(__int_as_float(rb0), p.s_out));
                if (fabsf(t1 - n1) > 0.49951171875f || !(fabsf(t1) < 4194304.f)) n1 = rintf(__fdiv_rn(__int_as_float(rb1), p.s_out));
                qo[2 * h] = static_cast<int>(fminf(fmaxf(__fadd_rn(n0, zo), -128.f), 127.f));
                qo[2 * h + 1] = static_cast<int>(fminf(fmaxf(__fadd_rn(n1, zo), -128.f), 127.f));
                if (p.binop == B200_BINOP_DIV) {  // 0 / 0: the reference's (int8_t)NaN is 0 on its x86 build
                    if (__int_as_float(rb0) != __int_as_float(rb0)) qo[2 * h] = 0;
                    if (__int_as_float(rb1) != __int_as_float(rb1)) qo[2 * h + 1] = 0;
                }
            }
            if (has_lut) {
                const uint32_t b0 = s_lut[qo[0] + 128], b1 = s_lut[qo[1] + 128], b2 = s_lut[qo[2] + 128], b3 = s_lut[qo[3] + 128];
                r[q] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
            } else {
                r[q] = pack4_i8(qo[0], qo[1], qo[2], qo[3]);
            }
        }
        out[i] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

// The residual add of a ResNet block (16 launches, 40 % of the batch-256 ResNet-50 step): the same float sequence as
// add_i8_kernel specialised for add / sub between two same-shape activations -- no run-time operator dispatch, the
// rounded quotient stays in its magic-number form (bits = kMagicI + rint(t)) and goes through two integer clamps into
// the post table or the byte pack instead of FADD + two FMNMX + F2I (F2I is an eighth-rate XU instruction on B200,
// tools/probes/pipe_rate.cu), and the range test of the magic rounding is done once on the host
// ((255 |s_a| + 255 |s_b|) / |s_out| < 2^22).  ~12 instead of ~20 instructions per element.
constexpr float kTieBand = 0.5f - 0.0001220703125f;  // 0.5 - 2^-13
template <bool HAS_LUT, bool SUB>
__global__ void __launch_bounds__(256) add_i8_fast_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                          uint4 *__restrict__ out, long long nvec, const AddArgs p)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __shared__ uint8_t s_lut[256];
    if (HAS_LUT)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = static_cast<uint8_t>(p.post_lut[i]);
    __syncthreads();
    const float off_a = -(kMagicF + 128.f + static_cast<float>(p.zp_a)), off_b = -(kMagicF + 128.f + static_cast<float>(p.zp_b));
    const float sb = SUB ? -p.s_b : p.s_b;  // a - b = a + (-(b * s_b)): negating the scale negates the rounded product exactly
    const uint64_t sa2 = f2_pack(p.s_a, p.s_a), sb2 = f2_pack(sb, sb), inv2 = f2_pack(p.inv_out, p.inv_out);
    const uint64_t mg2 = f2_pack(kMagicF, kMagicF), nmg2 = f2_pack(-kMagicF, -kMagicF);
    const int lut_base = static_cast<int>(smem_u32(s_lut));
    const int lut_lo = kMagicI - p.zp_out - 128 - lut_base;  // table address = bits - lut_lo, clamped to the table
    const int zp_m = p.zp_out - kMagicI;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 va = __ldg(a + i), vb = __ldg(b + i);
        const uint32_t wa[4] = {va.x ^ 0x80808080u, va.y ^ 0x80808080u, va.z ^ 0x80808080u, va.w ^ 0x80808080u};
        const uint32_t wb[4] = {vb.x ^ 0x80808080u, vb.y ^ 0x80808080u, vb.z ^ 0x80808080u, vb.w ^ 0x80808080u};
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int m[4], d[4], rb[4];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint64_t xa = f2_fma(bytes_to_f2(wa[q], 2 * h, off_a), sa2, 0ull);
                const uint64_t xb = f2_fma(bytes_to_f2(wb[q], 2 * h, off_b), sb2, 0ull);
                const uint64_t rr = f2_add(xa, xb);
                const uint64_t t2 = f2_fma(rr, inv2, 0ull);
                const uint64_t m2 = f2_add(t2, mg2);                                   // bits: kMagicI + rint(t)
                const uint64_t d2 = f2_fma(f2_add(m2, nmg2), f2_pack(-1.f, -1.f), t2);  // t - rint(t), exact
                f2_unpack_bits(m2, m[2 * h], m[2 * h + 1]);
                f2_unpack_bits(d2, d[2 * h], d[2 * h + 1]);
                f2_unpack_bits(rr, rb[2 * h], rb[2 * h + 1]);
            }
            // t is within |t| * 2^-23 of the IEEE quotient; where the result does not saturate either way (|t| < 512) that
            // is 2^-14, so rint(t) can only differ from rint(quotient) within 2^-13 of a half-integer: those elements
            // (0.02 %) take the real division.  One branch per four elements.
            bool any = false;
#pragma unroll
            for (int e = 0; e < 4; e++) any = any || fabsf(__int_as_float(d[e])) > kTieBand;
            if (any) {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (fabsf(__int_as_float(d[e])) > kTieBand)
                        m[e] = __float_as_int(__fadd_rn(rintf(__fdiv_rn(__int_as_float(rb[e]), p.s_out)), kMagicF));
            }
            if (HAS_LUT) {
                uint32_t bb[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int addr = min(max(m[e] - lut_lo, lut_base), lut_base + 255);
                    asm("ld.shared.u8 %0, [%1];" : "=r"(bb[e]) : "r"(addr));
                }
                r[q] = __byte_perm(__byte_perm(bb[0], bb[1], 0x0040), __byte_perm(bb[2], bb[3], 0x0040), 0x5410);
            } else {
                r[q] = pack4_sat_i8(m[0] + zp_m, m[1] + zp_m, m[2] + zp_m, m[3] + zp_m);
            }
        }
        out[i] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

__global__ void __launch_bounds__(256) add_f16_kernel(const uint4 *__restrict__ a,
                                                      const uint4 *__restrict__ b,
                                                      uint4 *__restrict__ out, long long nvec,
                                                      int act, int binop, const int b_period, const long long b_outer)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 va = __ldg(a + i), vb = __ldg(b + (b_period ? (b_outer ? i / b_outer * b_period : 0) + i % b_period : i));
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

static int binary_impl(int binop, int dtype, const void *a, const void *b, size_t b_count, size_t a_per_instance, void *out,
                       size_t count, float s_a, int zp_a, float s_b, int zp_b, float s_out, int zp_out,
                       const int8_t *post_lut, int act, void *stream);

extern "C" int b200_binary_bcast(int binop, int dtype, const void *a, const void *b, size_t b_count, void *out,
                                 size_t count, float s_a, int zp_a, float s_b, int zp_b, float s_out, int zp_out,
                                 const int8_t *post_lut, int act, void *stream)
{
    return binary_impl(binop, dtype, a, b, b_count, 0, out, count, s_a, zp_a, s_b, zp_b, s_out, zp_out, post_lut, act, stream);
}

extern "C" int b200_binary_bcast_nc(int binop, int dtype, const void *a, const void *b, size_t b_count, size_t a_per_instance,
                                    void *out, size_t count, float s_a, int zp_a, float s_b, int zp_b, float s_out, int zp_out,
                                    const int8_t *post_lut, int act, void *stream)
{
    return binary_impl(binop, dtype, a, b, b_count, a_per_instance, out, count, s_a, zp_a, s_b, zp_b, s_out, zp_out, post_lut, act,
                       stream);
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
    return b200_binary_bcast(binop, dtype, a, b, 0, out, count, s_a, zp_a, s_b, zp_b, s_out, zp_out, post_lut, act, stream);
}

// b_count = elements after which the second operand repeats (0 = same shape), a_per_instance = elements of the first operand
// that share one instance of that period (0 = one instance for the whole tensor: a constant; H * W * cp = one per image: a
// [N, C, 1, 1] activation against [N, C, H, W])
static int binary_impl(int binop, int dtype, const void *a, const void *b, size_t b_count, size_t a_per_instance, void *out,
                       size_t count, float s_a, int zp_a, float s_b, int zp_b, float s_out, int zp_out,
                       const int8_t *post_lut, int act, void *stream)
{
    if (binop < B200_BINOP_ADD || binop > B200_BINOP_DIV) {
        set_error("b200_binary: unknown op %d", binop);
        return B200_ERR_ARG;
    }
    const int vec = dtype == B200_I8 ? 16 : 8;
    if ((dtype != B200_I8 && dtype != B200_F16) || !a || !b || !out || count == 0 || count % vec ||
        !aligned16(a) || !aligned16(b) || !aligned16(out)) {
        set_error("b200_add: bad arguments (dtype=%d count=%zu)", dtype, count);
        return B200_ERR_ARG;
    }
    if (b_count % vec || b_count / vec >= (1u << 30)) {
        set_error("b200_binary_bcast: period %zu of the second operand is not a multiple of %d elements", b_count, vec);
        return B200_ERR_ARG;
    }
    const int b_period = static_cast<int>(b_count / vec);
    if (a_per_instance % vec || (a_per_instance && !b_period)) {
        set_error("b200_binary_bcast_nc: %zu elements per instance of the second operand's period %zu", a_per_instance, b_count);
        return B200_ERR_ARG;
    }
    const long long b_outer = static_cast<long long>(a_per_instance / vec);
    const long long nvec = static_cast<long long>(count / vec);
    if (dtype == B200_I8) {
        AddArgs p{s_a, s_b, s_out, zp_a, zp_b, zp_out, act, post_lut, binop, static_cast<float>(1.0 / static_cast<double>(s_out))};
        // add / sub of two activations with no fused clamp: the specialised kernel (its magic rounding needs |r / s_out| < 2^22)
        const double reach = (255.0 * fabs(static_cast<double>(s_a)) + 255.0 * fabs(static_cast<double>(s_b))) / fabs(static_cast<double>(s_out));
        if ((binop == B200_BINOP_ADD || binop == B200_BINOP_SUB) && b_period == 0 && act == B200_ACT_NONE && reach < 4194304.0 &&
            !getenv("SHL_B200_ADD_GENERIC")) {
            const dim3 g(ew_grid(nvec)), blk(256);
            cudaStream_t st = (cudaStream_t)stream;
            const uint4 *pa = static_cast<const uint4 *>(a), *pb = static_cast<const uint4 *>(b);
            uint4 *po = static_cast<uint4 *>(out);
            if (binop == B200_BINOP_ADD) {
                if (post_lut) launch_kernel(add_i8_fast_kernel<true, false>, g, blk, 0, st, pa, pb, po, nvec, p);
                else launch_kernel(add_i8_fast_kernel<false, false>, g, blk, 0, st, pa, pb, po, nvec, p);
            } else {
                if (post_lut) launch_kernel(add_i8_fast_kernel<true, true>, g, blk, 0, st, pa, pb, po, nvec, p);
                else launch_kernel(add_i8_fast_kernel<false, true>, g, blk, 0, st, pa, pb, po, nvec, p);
            }
            B200_LAUNCH_CHECK();
            return B200_OK;
        }
        launch_kernel(add_i8_kernel, dim3(ew_grid(nvec)), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const uint4 *>(a), static_cast<const uint4 *>(b), static_cast<uint4 *>(out),
            nvec, p, b_period, b_outer);
    } else {
        launch_kernel(add_f16_kernel, dim3(ew_grid(nvec)), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const uint4 *>(a), static_cast<const uint4 *>(b), static_cast<uint4 *>(out),
            nvec, act, binop, b_period, b_outer);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}
