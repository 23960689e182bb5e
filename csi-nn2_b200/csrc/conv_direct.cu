// conv_direct.cu -- int8 conv2d for a network's FIRST layer: few input channels (K = C*kh*kw
// <= 160), input still in the API's NCHW layout, output pixel-major.  With K = 27 (3x3x3) or 147
// (7x7x3) an im2col + GEMM round trip through HBM costs several times the layer's own traffic and
// the tensor core would multiply mostly zero padding, so this one layer shape is computed
// directly: one thread per output pixel gathers its K input bytes once (NCHW reads are coalesced
// across the warp: neighbouring threads own neighbouring pixels), packs them four to a word and
// runs dp4a against weight words held in shared memory (broadcast reads), 4 output channels per
// 128-bit shared load.  HBM traffic = input once + output once.
//
// k order = (ky, kx, c), the order b200_opt/quant.c packs conv weights in.  Epilogue = the
// contract of include/b200nn.h.  Replaces, for this shape, the im2col loop + 4x16 GEMM of
// shl_rvv_conv_im2col_gemm_int8 (source/thead_rvv/int8/convolution_gemm_int8.c:106-170).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

struct DirectArgs {
    int n, c, h, w, o, oh, ow, cp_out;
    int kh, kw, sh, sw, pt, pl, dh, dw;
    int kwords;  // ceil(K / 4)
    int ldw;     // weight row pitch in bytes (= ldk)
    const int8_t *in;
    const int8_t *wt;  // [O][ldw] bytes, k = (ky, kx, c)
    int8_t *out;
    int zp_in;
    EpiScalars ep;
};

constexpr int kDirectThreads = 128;

__global__ void __launch_bounds__(kDirectThreads) conv_direct_i8_kernel(const DirectArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    extern __shared__ uint32_t s_w[];  // [kwords][o4] words, o4 = O rounded up to 4
    const int o4 = (a.o + 3) & ~3;
    float *s_mu = reinterpret_cast<float *>(s_w + a.kwords * o4);
    float *s_ba = s_mu + o4;
    int *s_ib = reinterpret_cast<int *>(s_ba + o4);
    uint32_t *s_x = reinterpret_cast<uint32_t *>(s_ib + o4);  // [kwords][threads]: this thread's packed taps
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < a.kwords * o4; i += blockDim.x) {
        const int j = i / o4, o = i % o4;
        uint32_t wv = 0;
        if (o < a.o) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = j * 4 + e;
                const uint32_t byte = k < a.kh * a.kw * a.c ? static_cast<uint8_t>(a.wt[o * a.ldw + k]) : 0;
                wv |= byte << (8 * e);
            }
        }
        s_w[i] = wv;
    }
    for (int o = threadIdx.x; o < o4; o += blockDim.x) {
        s_mu[o] = o < a.o ? a.ep.mult[o] : 0.f;
        s_ba[o] = o < a.o ? a.ep.badd[o] : 0.f;
        s_ib[o] = o < a.o ? a.ep.ibias[o] : 0;
    }
    if (a.ep.post_lut != nullptr)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = static_cast<uint8_t>(a.ep.post_lut[i]);
    __syncthreads();
    const uint8_t *lut = a.ep.post_lut != nullptr ? s_lut : nullptr;

    const long long total = static_cast<long long>(a.n) * a.oh * a.ow;
    for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
         p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(p % a.ow);
        const int oy = static_cast<int>((p / a.ow) % a.oh);
        const int b = static_cast<int>(p / (static_cast<long long>(a.ow) * a.oh));
        // gather: K bytes of this pixel's receptive field, zp_in at padded taps, four to a word
        const int8_t *img = a.in + static_cast<long long>(b) * a.c * a.h * a.w;
        {
            int k = 0;
            uint32_t cur = 0;
            for (int ky = 0; ky < a.kh; ky++) {
                const int iy = oy * a.sh - a.pt + ky * a.dh;
                for (int kx = 0; kx < a.kw; kx++) {
                    const int ix = ox * a.sw - a.pl + kx * a.dw;
                    const bool ok = iy >= 0 && iy < a.h && ix >= 0 && ix < a.w;
                    for (int c = 0; c < a.c; c++, k++) {
                        const int v = ok ? img[(static_cast<long long>(c) * a.h + iy) * a.w + ix] : a.zp_in;
                        cur |= static_cast<uint32_t>(v & 0xFF) << (8 * (k & 3));
                        if ((k & 3) == 3) {
                            s_x[(k >> 2) * kDirectThreads + threadIdx.x] = cur;
                            cur = 0;
                        }
                    }
                }
            }
            if (k & 3) s_x[(k >> 2) * kDirectThreads + threadIdx.x] = cur;
        }
        int8_t *dst = a.out + p * a.cp_out;
        for (int ob = 0; ob < o4; ob += 16) {
            uint32_t pk[4] = {0, 0, 0, 0};
#pragma unroll
            for (int og = 0; og < 4; og++) {
                const int o = ob + og * 4;
                if (o >= o4) break;
                int acc[4];
                const int4 ib = *reinterpret_cast<const int4 *>(&s_ib[o]);
                acc[0] = ib.x, acc[1] = ib.y, acc[2] = ib.z, acc[3] = ib.w;
#pragma unroll 4
                for (int j = 0; j < a.kwords; j++) {
                    const int x = static_cast<int>(s_x[j * kDirectThreads + threadIdx.x]);
                    const uint4 wv = *reinterpret_cast<const uint4 *>(&s_w[j * o4 + o]);
                    acc[0] = __dp4a(x, static_cast<int>(wv.x), acc[0]);
                    acc[1] = __dp4a(x, static_cast<int>(wv.y), acc[1]);
                    acc[2] = __dp4a(x, static_cast<int>(wv.z), acc[2]);
                    acc[3] = __dp4a(x, static_cast<int>(wv.w), acc[3]);
                }
                const float4 mu = *reinterpret_cast<const float4 *>(&s_mu[o]);
                const float4 ba = *reinterpret_cast<const float4 *>(&s_ba[o]);
                const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, b4[4] = {ba.x, ba.y, ba.z, ba.w};
                int q[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    q[e] = __float2int_rn(fmaf(static_cast<float>(acc[e]), m4[e], b4[e])) + a.ep.zp_out;
                    if (a.ep.act != B200_ACT_NONE) q[e] = max(q[e], a.ep.zp_out);
                    if (a.ep.act == B200_ACT_RELU6) q[e] = min(q[e], a.ep.q6);
                }
                pk[og] = lut ? lut4_i8(q[0], q[1], q[2], q[3], lut) : pack4_sat_i8(q[0], q[1], q[2], q[3]);
            }
            if (ob < a.cp_out)
                *reinterpret_cast<uint4 *>(dst + ob) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// Compile-time (C, KH, KW) variant for the stems that matter (3x3x3, 7x7x3), dilation 1: the gather
// is fully unrolled (compile-time byte lanes, 32-bit index math, one predicated byte load per tap)
// and the packed taps stay in registers.
template <int C, int KH, int KW>
__global__ void __launch_bounds__(kDirectThreads) conv_direct_i8_sp_kernel(const DirectArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    constexpr int K = C * KH * KW;
    constexpr int KWORDS = (K + 3) / 4;
    constexpr bool MAGIC = K <= 128;  // |acc + ibias| < 2^22, see common.cuh
    extern __shared__ uint32_t s_w[];  // [KWORDS][o4]
    const int o4 = (a.o + 3) & ~3;
    float *s_mu = reinterpret_cast<float *>(s_w + KWORDS * o4);
    float *s_ba = s_mu + o4;
    int *s_ib = reinterpret_cast<int *>(s_ba + o4);
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < KWORDS * o4; i += blockDim.x) {
        const int j = i / o4, o = i % o4;
        uint32_t wv = 0;
        if (o < a.o) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = j * 4 + e;
                const uint32_t byte = k < K ? static_cast<uint8_t>(a.wt[o * a.ldw + k]) : 0;
                wv |= byte << (8 * e);
            }
        }
        s_w[i] = wv;
    }
    for (int o = threadIdx.x; o < o4; o += blockDim.x) {
        s_mu[o] = o < a.o ? a.ep.mult[o] : 0.f;
        s_ba[o] = o < a.o ? a.ep.badd[o] : 0.f;
        s_ib[o] = (o < a.o ? a.ep.ibias[o] : 0) + (MAGIC ? kMagicI : 0);
    }
    const bool has_lut = a.ep.post_lut != nullptr;
    if (has_lut)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = static_cast<uint8_t>(a.ep.post_lut[i]);
    __syncthreads();

    const int hw = a.h * a.w;
    const int opix = a.oh * a.ow;
    const int total = a.n * opix;  // < 2^31, host-checked
    const int zp_m = a.ep.zp_out - kMagicI;
    const uint32_t zpb = static_cast<uint32_t>(a.zp_in & 0xFF);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int b = p / opix;
        const int rem = p - b * opix;
        const int oy = rem / a.ow, ox = rem - oy * a.ow;
        const int8_t *img = a.in + static_cast<size_t>(b) * C * hw;
        uint32_t xw[KWORDS];
#pragma unroll
        for (int j = 0; j < KWORDS; j++) xw[j] = 0;
#pragma unroll
        for (int ky = 0; ky < KH; ky++) {
            const int iy = oy * a.sh - a.pt + ky;
            const bool yok = iy >= 0 && iy < a.h;
#pragma unroll
            for (int kx = 0; kx < KW; kx++) {
                const int ix = ox * a.sw - a.pl + kx;
                const bool ok = yok && ix >= 0 && ix < a.w;
                const int off = iy * a.w + ix;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int k = (ky * KW + kx) * C + c;
                    uint32_t v = zpb;
                    if (ok) v = static_cast<uint8_t>(img[c * hw + off]);
                    xw[k >> 2] |= v << (8 * (k & 3));
                }
            }
        }
        int8_t *dst = a.out + static_cast<size_t>(p) * a.cp_out;
        for (int ob = 0; ob < o4; ob += 16) {
            uint32_t pk[4] = {0, 0, 0, 0};
#pragma unroll
            for (int og = 0; og < 4; og++) {
                const int o = ob + og * 4;
                if (o >= o4) break;
                const int4 ib = *reinterpret_cast<const int4 *>(&s_ib[o]);
                int acc[4] = {ib.x, ib.y, ib.z, ib.w};
#pragma unroll
                for (int j = 0; j < KWORDS; j++) {
                    const uint4 wv = *reinterpret_cast<const uint4 *>(&s_w[j * o4 + o]);
                    acc[0] = __dp4a(static_cast<int>(xw[j]), static_cast<int>(wv.x), acc[0]);
                    acc[1] = __dp4a(static_cast<int>(xw[j]), static_cast<int>(wv.y), acc[1]);
                    acc[2] = __dp4a(static_cast<int>(xw[j]), static_cast<int>(wv.z), acc[2]);
                    acc[3] = __dp4a(static_cast<int>(xw[j]), static_cast<int>(wv.w), acc[3]);
                }
                const float4 mu = *reinterpret_cast<const float4 *>(&s_mu[o]);
                const float4 ba = *reinterpret_cast<const float4 *>(&s_ba[o]);
                const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, b4[4] = {ba.x, ba.y, ba.z, ba.w};
                int q[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float af = MAGIC ? magic_to_float(acc[e]) : static_cast<float>(acc[e]);
                    q[e] = magic_round(fmaf(af, m4[e], b4[e]), zp_m);
                    if (a.ep.act != B200_ACT_NONE) q[e] = max(q[e], a.ep.zp_out);
                    if (a.ep.act == B200_ACT_RELU6) q[e] = min(q[e], a.ep.q6);
                }
                pk[og] = has_lut ? lut4_i8(q[0], q[1], q[2], q[3], s_lut) : pack4_sat_i8(q[0], q[1], q[2], q[3]);
            }
            *reinterpret_cast<uint4 *>(dst + ob) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}


// fp16 first layer (few input channels, NCHW input): one thread per output pixel, f32 accumulation
// of all output channels in registers (packed f32x2 FMAs against f32 weights in shared memory),
// instead of an im2col round trip through HBM whose gather alone took 0.5 ms at batch 256.
// NO16 = output channels / 16, rounded up (<= 4).
// (CC, CKH, CKW) = compile-time shape for the stems that matter (3x3x3, 7x7x3: fully unrolled taps), 0 = runtime
template <int NO16, int CC, int CKH, int CKW>
__global__ void __launch_bounds__(kDirectThreads) conv_direct_f16_kernel(const DirectArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    constexpr int NO = NO16 * 16;
    extern __shared__ uint32_t s_w[];  // [K][NO] f32, then [NO] bias
    float *s_wf = reinterpret_cast<float *>(s_w);
    const int nc = CC ? CC : a.c, nkh = CKH ? CKH : a.kh, nkw = CKW ? CKW : a.kw;
    const int K = nc * nkh * nkw;
    float *s_b = s_wf + K * NO;
    const __half *wt = reinterpret_cast<const __half *>(a.wt);
    const int ldw = a.ldw / 2;  // halves
    for (int i = threadIdx.x; i < K * NO; i += blockDim.x) {
        const int k = i / NO, o = i % NO;
        s_wf[i] = o < a.o ? __half2float(wt[o * ldw + k]) : 0.f;
    }
    for (int o = threadIdx.x; o < NO; o += blockDim.x) s_b[o] = (o < a.o && a.ep.badd) ? a.ep.badd[o] : 0.f;
    __syncthreads();

    const __half *in = reinterpret_cast<const __half *>(a.in);
    __half *out = reinterpret_cast<__half *>(a.out);
    const int hw = a.h * a.w, opix = a.oh * a.ow;
    const int total = a.n * opix;  // < 2^31, host-checked
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int b = p / opix;
        const int rem = p - b * opix;
        const int oy = rem / a.ow, ox = rem - oy * a.ow;
        const __half *img = in + static_cast<size_t>(b) * nc * hw;
        uint64_t acc[NO / 2];
#pragma unroll
        for (int j = 0; j < NO / 2; j++) acc[j] = f2_pack(s_b[2 * j], s_b[2 * j + 1]);
        int k = 0;
#pragma unroll
        for (int ky = 0; ky < nkh; ky++) {
            const int iy = oy * a.sh - a.pt + ky * a.dh;
            const bool yok = iy >= 0 && iy < a.h;
#pragma unroll
            for (int kx = 0; kx < nkw; kx++) {
                const int ix = ox * a.sw - a.pl + kx * a.dw;
                const bool ok = yok && ix >= 0 && ix < a.w;
#pragma unroll
                for (int c = 0; c < nc; c++, k++) {
                    const float x = ok ? __half2float(__ldg(img + c * hw + iy * a.w + ix)) : 0.f;
                    const uint64_t x2 = f2_pack(x, x);
                    const ulonglong2 *wrow = reinterpret_cast<const ulonglong2 *>(s_wf + k * NO);
#pragma unroll
                    for (int j = 0; j < NO / 4; j++) {
                        const ulonglong2 w4 = wrow[j];  // four f32 weights, broadcast to the warp
                        acc[2 * j] = f2_fma(x2, w4.x, acc[2 * j]);
                        acc[2 * j + 1] = f2_fma(x2, w4.y, acc[2 * j + 1]);
                    }
                }
            }
        }
        __half *dst = out + static_cast<size_t>(p) * a.cp_out;
#pragma unroll
        for (int v = 0; v < NO / 8; v++) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int b0, b1;
                f2_unpack_bits(acc[v * 4 + q], b0, b1);
                const int o = v * 8 + q * 2;
                float f0 = act_f(__int_as_float(b0), a.ep.act), f1 = act_f(__int_as_float(b1), a.ep.act);
                f0 = o < a.o ? f0 : 0.f, f1 = o + 1 < a.o ? f1 : 0.f;
                const __half2 hv = __floats2half2_rn(f0, f1);
                pk[q] = *reinterpret_cast<const uint32_t *>(&hv);
            }
            if (v * 8 < a.cp_out) *reinterpret_cast<uint4 *>(dst + v * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

}  // namespace b200

using namespace b200;

int b200_conv_stem_tc_launch(const b200_conv_direct_desc *d, void *stream);      // conv_stem_tc.cu
int b200_conv_stem_f16_tc_launch(const b200_conv_direct_desc *d, void *stream);  // conv_stem_f16_tc.cu

extern "C" int b200_conv2d_direct(const b200_conv_direct_desc *d, void *stream)
{
    if (!d || !d->in || !d->wt || !d->out || !d->ep.mult || !d->ep.badd || !d->ep.ibias) {
        set_error("b200_conv2d_direct: null descriptor field");
        return B200_ERR_ARG;
    }
    const int K = d->c * d->kh * d->kw;
    if (d->n <= 0 || d->c <= 0 || d->o <= 0 || K > 160 || d->o > 256 || d->cp_out < d->o || d->cp_out % 16 ||
        d->stride_h < 1 || d->stride_w < 1 || d->dil_h < 1 || d->dil_w < 1 || d->ldw < K) {
        set_error("b200_conv2d_direct: unsupported shape (C*kh*kw=%d must be <= 160, O=%d <= 256)", K, d->o);
        return B200_ERR_UNSUPPORTED;
    }
    // the 3-channel stems run as an implicit GEMM on the tensor core (conv_stem_tc.cu)
    if (!getenv("SHL_B200_NO_STEM_TC")) {
        const int rc = b200_conv_stem_tc_launch(d, stream);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    DirectArgs a;
    a.n = d->n, a.c = d->c, a.h = d->h, a.w = d->w, a.o = d->o, a.oh = d->oh, a.ow = d->ow, a.cp_out = d->cp_out;
    a.kh = d->kh, a.kw = d->kw, a.sh = d->stride_h, a.sw = d->stride_w, a.pt = d->pad_top, a.pl = d->pad_left;
    a.dh = d->dil_h, a.dw = d->dil_w, a.kwords = (K + 3) / 4, a.ldw = d->ldw;
    a.in = static_cast<const int8_t *>(d->in), a.wt = static_cast<const int8_t *>(d->wt);
    a.out = static_cast<int8_t *>(d->out), a.zp_in = d->zp_in, a.ep = make_epi(d->ep);
    const int o4 = (d->o + 3) & ~3;
    const size_t smem = static_cast<size_t>(a.kwords) * o4 * 4 + static_cast<size_t>(o4) * 12 +
                        static_cast<size_t>(a.kwords) * kDirectThreads * 4;
    const long long total = static_cast<long long>(d->n) * d->oh * d->ow;
    long long g = (total + 127) / 128;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    const int grid = static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
    const bool unit_dil = d->dil_h == 1 && d->dil_w == 1;
    const size_t smem_sp = static_cast<size_t>(a.kwords) * o4 * 4 + static_cast<size_t>(o4) * 12;
    const bool special = unit_dil && d->c == 3 && d->kh == d->kw && (d->kh == 3 || d->kh == 7);
    // the same budget b200_opt/ops.c:conv_goes_direct applies before choosing this kernel over im2col + GEMM
    if ((special ? smem_sp : smem) > 48 * 1024) {
        set_error("b200_conv2d_direct: weights + taps (%zu bytes) exceed the shared-memory budget", special ? smem_sp : smem);
        return B200_ERR_UNSUPPORTED;
    }
    if (total >= (1ll << 31)) {
        set_error("b200_conv2d_direct: %lld output pixels exceed the 32-bit pixel index", total);
        return B200_ERR_UNSUPPORTED;
    }
    if (unit_dil && d->c == 3 && d->kh == 3 && d->kw == 3)
        launch_kernel(conv_direct_i8_sp_kernel<3, 3, 3>, dim3(grid), dim3(kDirectThreads), smem_sp, (cudaStream_t)stream, a);
    else if (unit_dil && d->c == 3 && d->kh == 7 && d->kw == 7)
        launch_kernel(conv_direct_i8_sp_kernel<3, 7, 7>, dim3(grid), dim3(kDirectThreads), smem_sp, (cudaStream_t)stream, a);
    else
        launch_kernel(conv_direct_i8_kernel, dim3(grid), dim3(kDirectThreads), smem, (cudaStream_t)stream, a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// fp16 twin of b200_conv2d_direct (same descriptor, ldw in bytes): first layers with C*kh*kw <= 160
// and at most 64 output channels
extern "C" int b200_conv2d_direct_f16(const b200_conv_direct_desc *d, void *stream)
{
    if (!d || !d->in || !d->wt || !d->out) {
        set_error("b200_conv2d_direct_f16: null descriptor field");
        return B200_ERR_ARG;
    }
    const int K = d->c * d->kh * d->kw;
    const int no16 = (d->o + 15) / 16;
    const long long total = static_cast<long long>(d->n) * d->oh * d->ow;
    if (d->n <= 0 || d->c <= 0 || d->o <= 0 || K > 160 || no16 > 4 || d->cp_out < d->o || d->cp_out % 8 ||
        d->stride_h < 1 || d->stride_w < 1 || d->dil_h < 1 || d->dil_w < 1 || d->ldw < 2 * K || total >= (1ll << 31) ||
        static_cast<long long>(d->n) * d->c * d->h * d->w >= (1ll << 31)) {
        set_error("b200_conv2d_direct_f16: unsupported shape (C*kh*kw=%d must be <= 160, O=%d <= 64)", K, d->o);
        return B200_ERR_UNSUPPORTED;
    }
    // the 3-channel 3x3 stem runs as an implicit GEMM on the tensor core (conv_stem_f16_tc.cu)
    if (!getenv("SHL_B200_NO_STEM_TC")) {
        const int rc = b200_conv_stem_f16_tc_launch(d, stream);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    DirectArgs a;
    a.n = d->n, a.c = d->c, a.h = d->h, a.w = d->w, a.o = d->o, a.oh = d->oh, a.ow = d->ow, a.cp_out = d->cp_out;
    a.kh = d->kh, a.kw = d->kw, a.sh = d->stride_h, a.sw = d->stride_w, a.pt = d->pad_top, a.pl = d->pad_left;
    a.dh = d->dil_h, a.dw = d->dil_w, a.kwords = 0, a.ldw = d->ldw;
    a.in = static_cast<const int8_t *>(d->in), a.wt = static_cast<const int8_t *>(d->wt);
    a.out = static_cast<int8_t *>(d->out), a.zp_in = 0, a.ep = make_epi(d->ep);
    const size_t smem = (static_cast<size_t>(K) * no16 * 16 + no16 * 16) * sizeof(float);
    long long g = (total + kDirectThreads - 1) / kDirectThreads;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    const int grid = static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
    cudaStream_t s = (cudaStream_t)stream;
    const bool s3 = d->c == 3 && d->kh == 3 && d->kw == 3, s7 = d->c == 3 && d->kh == 7 && d->kw == 7;
#define B200_F16_DIRECT(N)                                                                                        \
    case N:                                                                                                       \
        if (s3)                                                                                                   \
            launch_kernel(conv_direct_f16_kernel<N, 3, 3, 3>, dim3(grid), dim3(kDirectThreads), smem, s, a);      \
        else if (s7)                                                                                              \
            launch_kernel(conv_direct_f16_kernel<N, 3, 7, 7>, dim3(grid), dim3(kDirectThreads), smem, s, a);      \
        else                                                                                                      \
            launch_kernel(conv_direct_f16_kernel<N, 0, 0, 0>, dim3(grid), dim3(kDirectThreads), smem, s, a);      \
        break;
    switch (no16) {
        B200_F16_DIRECT(1)
        B200_F16_DIRECT(2)
        B200_F16_DIRECT(3)
        default:
            B200_F16_DIRECT(4)
    }
#undef B200_F16_DIRECT
    B200_LAUNCH_CHECK();
    return B200_OK;
}
