// dwconv.cu -- depthwise conv2d on pixel-major tensors (HBM-bound stencil).
//
// One thread owns 16 bytes of channels (16 int8 / 8 fp16) of kTW consecutive output pixels of
// one row, so every global access is a 128-bit vector and a warp covers 512 contiguous bytes of
// the channel axis (or several pixels when C is small).  int8 products use dp4a against
// "expanded" weights (one weight byte per 32-bit word, the other three zero), which gives the
// exact per-channel int32 product without unpacking the activations; accumulation is int32 and
// the epilogue is the contract of include/b200nn.h.  fp16 accumulates in fp32 like the
// reference (source/reference/convolution.c:206-269).
//
// Replaces shl_rvv_dwconv3x3s1_int8 / s2 (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31)
// and the fp16 twins; any kernel size / stride / dilation is covered by the same kernel.
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

constexpr int kTW = 4;  // output pixels per thread along W

struct DwArgs {
    int n, c, cp, h, w, oh, ow;
    int kh, kw, sh, sw, pt, pl, dh, dw;
    const void *in;
    const void *wt;  // int8: expanded [kh*kw][cp] words ; fp16: [kh*kw][cp] halves
    void *out;
    int zp_in;
    EpiScalars ep;
    const int32_t *wzp;  // int8 asymmetric weights: per-channel weight zero points, else null
};

// ASYM: weights with a zero point -- the window sum S of the (zero-point padded) taps is accumulated per channel
// beside the products and acc - zw * S enters the epilogue (ibias carries the constant part, include/b200nn.h)
template <bool ASYM>
__global__ void __launch_bounds__(128) dwconv_i8_kernel(const DwArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __shared__ int8_t s_lut[256];
    if (a.ep.post_lut != nullptr)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = a.ep.post_lut[i];
    __syncthreads();

    const int chunks = a.cp / 16;
    const int xgroups = (a.ow + kTW - 1) / kTW;
    const long long total = static_cast<long long>(a.n) * a.oh * xgroups * chunks;
    const int8_t *in = static_cast<const int8_t *>(a.in);
    const uint32_t *wexp = static_cast<const uint32_t *>(a.wt);
    int8_t *out = static_cast<int8_t *>(a.out);
    const uint32_t padw = 0x01010101u * static_cast<uint32_t>(a.zp_in & 0xFF);

    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % chunks);
        const int xg = static_cast<int>((i / chunks) % xgroups);
        const int oy = static_cast<int>((i / (static_cast<long long>(chunks) * xgroups)) % a.oh);
        const int b = static_cast<int>(i / (static_cast<long long>(chunks) * xgroups * a.oh));
        const int c0 = ch * 16;
        const int ox0 = xg * kTW;

        int acc[kTW][16];
        int xs[ASYM ? kTW : 1][ASYM ? 16 : 1];
#pragma unroll
        for (int p = 0; p < kTW; p++)
#pragma unroll
            for (int j = 0; j < 16; j++) {
                acc[p][j] = 0;
                if (ASYM) xs[p][j] = 0;
            }

        for (int ky = 0; ky < a.kh; ky++) {
            const int iy = oy * a.sh - a.pt + ky * a.dh;
            const bool yok = iy >= 0 && iy < a.h;
            for (int kx = 0; kx < a.kw; kx++) {
                // 16 expanded weight words for this tap and these 16 channels
                const uint4 *wp = reinterpret_cast<const uint4 *>(
                    wexp + (static_cast<long long>(ky * a.kw + kx) * a.cp + c0));
                uint32_t wv[16];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 t = __ldg(wp + q);
                    wv[q * 4 + 0] = t.x, wv[q * 4 + 1] = t.y, wv[q * 4 + 2] = t.z, wv[q * 4 + 3] = t.w;
                }
#pragma unroll
                for (int p = 0; p < kTW; p++) {
                    const int ix = (ox0 + p) * a.sw - a.pl + kx * a.dw;
                    uint4 x = make_uint4(padw, padw, padw, padw);
                    if (yok && ix >= 0 && ix < a.w && ox0 + p < a.ow)
                        x = __ldg(reinterpret_cast<const uint4 *>(
                            in + ((static_cast<long long>(b) * a.h + iy) * a.w + ix) * a.cp + c0));
                    const uint32_t xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            acc[p][q * 4 + e] = __dp4a(static_cast<int>(xw[q]),
                                                       static_cast<int>(wv[q * 4 + e]),
                                                       acc[p][q * 4 + e]);
                            if (ASYM) xs[p][q * 4 + e] = __dp4a(static_cast<int>(xw[q]), 1 << (8 * e), xs[p][q * 4 + e]);
                        }
                }
            }
        }

        // epilogue: per-channel parameters once per thread, reused for the kTW pixels
        float mu[16], ba[16];
        int ib[16];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 m4 = __ldg(reinterpret_cast<const float4 *>(a.ep.mult + c0) + q);
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.ep.badd + c0) + q);
            const int4 i4 = __ldg(reinterpret_cast<const int4 *>(a.ep.ibias + c0) + q);
            mu[q * 4] = m4.x, mu[q * 4 + 1] = m4.y, mu[q * 4 + 2] = m4.z, mu[q * 4 + 3] = m4.w;
            ba[q * 4] = b4.x, ba[q * 4 + 1] = b4.y, ba[q * 4 + 2] = b4.z, ba[q * 4 + 3] = b4.w;
            ib[q * 4] = i4.x, ib[q * 4 + 1] = i4.y, ib[q * 4 + 2] = i4.z, ib[q * 4 + 3] = i4.w;
        }
        if (ASYM) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int4 z4 = __ldg(reinterpret_cast<const int4 *>(a.wzp + c0) + q);
                const int z[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
                for (int p = 0; p < kTW; p++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[p][q * 4 + e] -= z[e] * xs[p][q * 4 + e];
            }
        }
#pragma unroll
        for (int p = 0; p < kTW; p++) {
            if (ox0 + p >= a.ow) break;
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int v[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int j = q * 4 + e;
                    int qv = requant_i8(acc[p][j] + ib[j], mu[j], ba[j], a.ep.zp_out, a.ep.act,
                                        a.ep.q6);
                    if (a.ep.post_lut != nullptr) qv = s_lut[qv + 128];
                    v[e] = c0 + j < a.c ? qv : 0;
                }
                pk[q] = pack4_i8(v[0], v[1], v[2], v[3]);
            }
            *reinterpret_cast<uint4 *>(
                out + ((static_cast<long long>(b) * a.oh + oy) * a.ow + ox0 + p) * a.cp + c0) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

__global__ void __launch_bounds__(128) dwconv_f16_kernel(const DwArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    const int chunks = a.cp / 8;
    const int xgroups = (a.ow + kTW - 1) / kTW;
    const long long total = static_cast<long long>(a.n) * a.oh * xgroups * chunks;
    const __half *in = static_cast<const __half *>(a.in);
    const __half *wt = static_cast<const __half *>(a.wt);
    __half *out = static_cast<__half *>(a.out);

    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % chunks);
        const int xg = static_cast<int>((i / chunks) % xgroups);
        const int oy = static_cast<int>((i / (static_cast<long long>(chunks) * xgroups)) % a.oh);
        const int b = static_cast<int>(i / (static_cast<long long>(chunks) * xgroups * a.oh));
        const int c0 = ch * 8;
        const int ox0 = xg * kTW;

        float acc[kTW][8];
#pragma unroll
        for (int p = 0; p < kTW; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[p][j] = 0.f;

        for (int ky = 0; ky < a.kh; ky++) {
            const int iy = oy * a.sh - a.pt + ky * a.dh;
            const bool yok = iy >= 0 && iy < a.h;
            for (int kx = 0; kx < a.kw; kx++) {
                const uint4 wraw = __ldg(reinterpret_cast<const uint4 *>(
                    wt + static_cast<long long>(ky * a.kw + kx) * a.cp + c0));
                const __half2 *wh = reinterpret_cast<const __half2 *>(&wraw);
                float wf[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float2 f = __half22float2(wh[q]);
                    wf[q * 2] = f.x, wf[q * 2 + 1] = f.y;
                }
#pragma unroll
                for (int p = 0; p < kTW; p++) {
                    const int ix = (ox0 + p) * a.sw - a.pl + kx * a.dw;
                    if (yok && ix >= 0 && ix < a.w && ox0 + p < a.ow) {
                        const uint4 xraw = __ldg(reinterpret_cast<const uint4 *>(
                            in + ((static_cast<long long>(b) * a.h + iy) * a.w + ix) * a.cp + c0));
                        const __half2 *xh = reinterpret_cast<const __half2 *>(&xraw);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float2 f = __half22float2(xh[q]);
                            acc[p][q * 2] = fmaf(wf[q * 2], f.x, acc[p][q * 2]);
                            acc[p][q * 2 + 1] = fmaf(wf[q * 2 + 1], f.y, acc[p][q * 2 + 1]);
                        }
                    }
                }
            }
        }
        float ba[8];
#pragma unroll
        for (int j = 0; j < 8; j++) ba[j] = a.ep.badd ? __ldg(a.ep.badd + c0 + j) : 0.f;
#pragma unroll
        for (int p = 0; p < kTW; p++) {
            if (ox0 + p >= a.ow) break;
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float f0 = act_f(acc[p][q * 2] + ba[q * 2], a.ep.act);
                float f1 = act_f(acc[p][q * 2 + 1] + ba[q * 2 + 1], a.ep.act);
                f0 = c0 + q * 2 < a.c ? f0 : 0.f;
                f1 = c0 + q * 2 + 1 < a.c ? f1 : 0.f;
                __half2 hv = __floats2half2_rn(f0, f1);
                pk[q] = *reinterpret_cast<uint32_t *>(&hv);
            }
            *reinterpret_cast<uint4 *>(
                out + ((static_cast<long long>(b) * a.oh + oy) * a.ow + ox0 + p) * a.cp + c0) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}


// fp16 depthwise 3x3, stride 1 / 2, dilation 1: one thread owns FOUR channels of TWO adjacent
// output columns and slides down a band of output rows.  Every input row is loaded once per thread
// (4 / 5 pixels x 8 bytes), converted to f32 once, and feeds the three output rows it belongs to
// through rotating accumulator sets (f32 accumulation like the reference, packed f32x2 FMAs), so
// vertical taps are reused from registers and horizontal ones from the same loads -- the generic
// kernel above reloads every tap.  Weights tap-major [9][cp] halves; bias seeds the accumulators.
template <int S>
__global__ void __launch_bounds__(128, 5) dwconv3x3_f16_kernel(const DwArgs a, int band_rows, int ybands)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    constexpr int NX = S + 3;  // input columns per row for two output columns
    const int quads = a.cp / 4;
    const int xpairs = (a.ow + 1) / 2;
    const long long total = static_cast<long long>(a.n) * ybands * xpairs * quads;
    const __half *in = static_cast<const __half *>(a.in);
    const __half *wt = static_cast<const __half *>(a.wt);
    __half *out = static_cast<__half *>(a.out);
    const long long in_row = static_cast<long long>(a.w) * a.cp, out_row = static_cast<long long>(a.ow) * a.cp;  // halves
    const bool clamp_lo = a.ep.act != B200_ACT_NONE, clamp_hi = a.ep.act == B200_ACT_RELU6;
    const __half2 zero2 = __floats2half2_rn(0.f, 0.f), six2 = __floats2half2_rn(6.f, 6.f);

    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cq = static_cast<int>(i % quads);
        long long rest = i / quads;
        const int xp = static_cast<int>(rest % xpairs);
        rest /= xpairs;
        const int yb = static_cast<int>(rest % ybands);
        const int b = static_cast<int>(rest / ybands);
        const int c0 = cq * 4;
        const int ox0 = xp * 2;
        const int oy0 = yb * band_rows;
        const int rows = min(band_rows, a.oh - oy0);

        // weights of the 9 taps and the bias of these four channels, as f32x2 pairs
        uint64_t w[3][3][2], seed[2];
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(wt + static_cast<size_t>(t) * a.cp + c0));
            const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
            w[t / 3][t % 3][0] = f2_pack(lo.x, lo.y), w[t / 3][t % 3][1] = f2_pack(hi.x, hi.y);
        }
        {
            const float b0 = a.ep.badd ? __ldg(a.ep.badd + c0) : 0.f, b1 = a.ep.badd ? __ldg(a.ep.badd + c0 + 1) : 0.f;
            const float b2 = a.ep.badd ? __ldg(a.ep.badd + c0 + 2) : 0.f, b3 = a.ep.badd ? __ldg(a.ep.badd + c0 + 3) : 0.f;
            seed[0] = f2_pack(b0, b1), seed[1] = f2_pack(b2, b3);
        }
        // Addressing is hoisted out of the row loop (it used to cost three times the FMAs): one input
        // pointer stepping a row at a time, column validity as a bit mask, rows tested against one bound.
        const int ix0 = ox0 * S - a.pl;
        int xmask = 0;
#pragma unroll
        for (int p = 0; p < NX; p++) xmask |= (ix0 + p >= 0 && ix0 + p < a.w) ? 1 << p : 0;
        const int iy0 = oy0 * S - a.pt;
        const int iy_max = min(a.h - 1, (oy0 + rows - 1) * S - a.pt + 2);
        // row iy0, column ix0 of this image (dereferenced only where both are inside it)
        const __half *rp = in + (static_cast<long long>(b) * a.h + iy0) * in_row + static_cast<long long>(ix0) * a.cp + c0;
        int iy = iy0;

        // one input row: NX pixels x 4 channels (zero outside the image: fp16 padding is 0).  The raw
        // halves of row r + 1 are requested before row r is used, so every thread keeps two rows of
        // loads in flight.
        uint2 raw[NX];
        auto fetch = [&]() {
            const bool yok = iy >= 0 && iy <= iy_max;
#pragma unroll
            for (int p = 0; p < NX; p++) {
                raw[p] = make_uint2(0u, 0u);
                if (yok && ((xmask >> p) & 1)) raw[p] = __ldg(reinterpret_cast<const uint2 *>(rp + p * a.cp));
            }
            rp += in_row, iy++;
        };
        // converts the fetched row to f32x2 pairs and requests the next one
        auto load_row = [&](uint64_t (&x)[NX][2]) {
#pragma unroll
            for (int p = 0; p < NX; p++) {
                const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw[p].x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw[p].y));
                x[p][0] = f2_pack(lo.x, lo.y), x[p][1] = f2_pack(hi.x, hi.y);
            }
            fetch();
        };
        // acc[col][pair] += row (x) kernel row ky
        auto fma_row = [&](uint64_t (&acc)[2][2], const uint64_t (&x)[NX][2], int ky) {
#pragma unroll
            for (int col = 0; col < 2; col++)
#pragma unroll
                for (int kx = 0; kx < 3; kx++)
#pragma unroll
                    for (int h = 0; h < 2; h++) acc[col][h] = f2_fma(x[col * S + kx][h], w[ky][kx][h], acc[col][h]);
        };
        auto start = [&](uint64_t (&acc)[2][2], const uint64_t (&x)[NX][2]) {
#pragma unroll
            for (int col = 0; col < 2; col++)
#pragma unroll
                for (int h = 0; h < 2; h++) acc[col][h] = seed[h];
            fma_row(acc, x, 0);
        };
        // padded channels are written as zeros; relu / relu6 clamp the rounded halves (0 and 6 are exact)
        const uint32_t m01 = (c0 < a.c ? 0xFFFFu : 0u) | (c0 + 1 < a.c ? 0xFFFF0000u : 0u);
        const uint32_t m23 = (c0 + 2 < a.c ? 0xFFFFu : 0u) | (c0 + 3 < a.c ? 0xFFFF0000u : 0u);
        const bool col1 = ox0 + 1 < a.ow;
        __half *op = out + ((static_cast<long long>(b) * a.oh + oy0) * a.ow + ox0) * a.cp + c0;
        auto store = [&](const uint64_t (&acc)[2][2]) {
#pragma unroll
            for (int col = 0; col < 2; col++) {
                if (col == 0 || col1) {
                    int b0, b1, b2, b3;
                    f2_unpack_bits(acc[col][0], b0, b1);
                    f2_unpack_bits(acc[col][1], b2, b3);
                    __half2 h01 = __floats2half2_rn(__int_as_float(b0), __int_as_float(b1));
                    __half2 h23 = __floats2half2_rn(__int_as_float(b2), __int_as_float(b3));
                    if (clamp_lo) h01 = __hmax2(h01, zero2), h23 = __hmax2(h23, zero2);
                    if (clamp_hi) h01 = __hmin2(h01, six2), h23 = __hmin2(h23, six2);
                    *reinterpret_cast<uint2 *>(op + col * a.cp) =
                        make_uint2(*reinterpret_cast<const uint32_t *>(&h01) & m01, *reinterpret_cast<const uint32_t *>(&h23) & m23);
                }
            }
            op += out_row;
        };

        uint64_t x[NX][2], accA[2][2], accB[2][2], accC[2][2];
        fetch();
        if (S == 1) {
            // input row r (image row iy0 + r) feeds output rows r (ky 0), r - 1 (ky 1), r - 2 (ky 2, completes it)
            load_row(x);
            start(accA, x);
            load_row(x);
            fma_row(accA, x, 1), start(accB, x);
            for (int y = 0;; y += 3) {
                load_row(x);
                fma_row(accA, x, 2), fma_row(accB, x, 1), start(accC, x);
                store(accA);
                if (y + 1 >= rows) break;
                load_row(x);
                fma_row(accB, x, 2), fma_row(accC, x, 1), start(accA, x);
                store(accB);
                if (y + 2 >= rows) break;
                load_row(x);
                fma_row(accC, x, 2), fma_row(accA, x, 1), start(accB, x);
                store(accC);
                if (y + 3 >= rows) break;
            }
        } else {
            // output row y reads input rows 2y (ky 0), 2y + 1 (ky 1), 2y + 2 (ky 2 = ky 0 of row y + 1)
            load_row(x);
            start(accA, x);
            for (int y = 0;; y += 2) {
                load_row(x);
                fma_row(accA, x, 1);
                load_row(x);
                fma_row(accA, x, 2), start(accB, x);
                store(accA);
                if (y + 1 >= rows) break;
                load_row(x);
                fma_row(accB, x, 1);
                load_row(x);
                fma_row(accB, x, 2), start(accA, x);
                store(accB);
                if (y + 2 >= rows) break;
            }
        }
    }
}

}  // namespace b200

using namespace b200;

int b200_dwconv3x3_tma_launch(const b200_dwconv_desc *d, const void *wrow, void *stream);  // dwconv3x3_tma.cu
int b200_dwconv3x3_umma_launch(const b200_dwconv_desc *d, const void *wrow, void *stream, int *handled);  // dwconv3x3_umma.cu
int b200_dwconv3x3_umma128_launch(const b200_dwconv_desc *d, const void *wrow, void *stream, int *handled);  // dwconv3x3_umma128.cu
int b200_dwconv3x3_imma_launch(const b200_dwconv_desc *d, const void *wrow, void *stream, int *handled);  // dwconv3x3_imma.cu

extern "C" int b200_dwconv2d(const b200_dwconv_desc *d, void *stream)
{
    if (!d || !d->in || !d->wt || !d->out) {
        set_error("b200_dwconv2d: null descriptor field");
        return B200_ERR_ARG;
    }
    if (d->dtype != B200_I8 && d->dtype != B200_F16) {
        set_error("b200_dwconv2d: dtype %d unsupported", d->dtype);
        return B200_ERR_UNSUPPORTED;
    }
    const int eb = d->dtype == B200_I8 ? 1 : 2;
    if (d->n <= 0 || d->c <= 0 || d->cp < d->c || (d->cp * eb) % 16 || d->kh <= 0 || d->kw <= 0 ||
        d->stride_h <= 0 || d->stride_w <= 0 || d->dil_h <= 0 || d->dil_w <= 0 || d->oh <= 0 ||
        d->ow <= 0 || (d->dtype == B200_I8 && (!d->ep.mult || !d->ep.badd || !d->ep.ibias))) {
        set_error("b200_dwconv2d: bad descriptor (c=%d cp=%d k=%dx%d)", d->c, d->cp, d->kh, d->kw);
        return B200_ERR_ARG;
    }
    if (d->dtype == B200_I8 && d->wt_row3 && d->kh == 3 && d->kw == 3 && d->dil_h == 1 && d->dil_w == 1 &&
        !getenv("SHL_B200_DW_GENERIC") && !d->w_zp &&
        d->stride_h == d->stride_w && (d->stride_h == 1 || d->stride_h == 2) &&
        // the TMA kernel folds zero-point padding into per-class accumulator seeds: at most one
        // padded row / column on each side of any output's 3x3 window
        d->pad_top <= 1 && d->pad_left <= 1 && (d->oh - 1) * d->stride_h - d->pad_top + 2 <= d->h &&
        (d->ow - 1) * d->stride_w - d->pad_left + 2 <= d->w) {
        // stride 1 with "same" padding: the taps are accumulated on the tensor cores
        int handled = 0;
        const char *tc = getenv("SHL_B200_DW_UMMA");
        if (tc && atoi(tc) == 2) {
            const int rc2 = b200_dwconv3x3_umma128_launch(d, d->wt_row3, stream, &handled);
            if (rc2 || handled) return rc2;
        }
        const int rc = b200_dwconv3x3_umma_launch(d, d->wt_row3, stream, &handled);
        if (rc || handled) return rc;
        // the default: taps on the warp-level tensor path (diagonal-B IMMA); maps / channel counts it leaves alone go on
        const int rc3 = b200_dwconv3x3_imma_launch(d, d->wt_row3, stream, &handled);
        if (rc3 || handled) return rc3;
        return b200_dwconv3x3_tma_launch(d, d->wt_row3, stream);
    }
    DwArgs a;
    a.wzp = nullptr;
    a.n = d->n, a.c = d->c, a.cp = d->cp, a.h = d->h, a.w = d->w, a.oh = d->oh, a.ow = d->ow;
    a.kh = d->kh, a.kw = d->kw, a.sh = d->stride_h, a.sw = d->stride_w;
    if (d->dtype == B200_F16 && d->kh == 3 && d->kw == 3 && d->dil_h == 1 && d->dil_w == 1 &&
        d->stride_h == d->stride_w && (d->stride_h == 1 || d->stride_h == 2) && !getenv("SHL_B200_DW_GENERIC")) {
        // register-sliding fp16 3x3: bands of output rows per thread, enough threads for ~8 waves
        a.pt = d->pad_top, a.pl = d->pad_left, a.dh = 1, a.dw = 1;
        a.in = d->in, a.wt = d->wt, a.out = d->out, a.zp_in = 0;
        a.ep = make_epi(d->ep);
        const long long cols = static_cast<long long>(d->n) * ((d->ow + 1) / 2) * (d->cp / 4);
        int ybands = 1;
        while (ybands < d->oh && cols * ybands < static_cast<long long>(sm_count()) * 128 * 16 && d->oh / (ybands * 2) >= 4)
            ybands *= 2;
        const int band_rows = (d->oh + ybands - 1) / ybands;
        ybands = (d->oh + band_rows - 1) / band_rows;
        const long long total = cols * ybands;
        long long g = (total + 127) / 128;
        const long long cap = static_cast<long long>(sm_count()) * 32;
        const int grid = static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
        if (d->stride_h == 1)
            launch_kernel(dwconv3x3_f16_kernel<1>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a, band_rows, ybands);
        else
            launch_kernel(dwconv3x3_f16_kernel<2>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a, band_rows, ybands);
        B200_LAUNCH_CHECK();
        return B200_OK;
    }
    a.pt = d->pad_top, a.pl = d->pad_left, a.dh = d->dil_h, a.dw = d->dil_w;
    a.in = d->in, a.wt = d->wt, a.out = d->out, a.zp_in = d->zp_in;
    a.ep = make_epi(d->ep);
    a.wzp = d->dtype == B200_I8 ? d->w_zp : nullptr;
    const int vec = 16 / eb;
    const long long total = static_cast<long long>(d->n) * d->oh * ((d->ow + kTW - 1) / kTW) *
                            (d->cp / vec);
    long long g = (total + 127) / 128;
    const long long cap = static_cast<long long>(sm_count()) * 32;
    const int grid = static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
    if (d->dtype == B200_I8 && a.wzp)
        launch_kernel(dwconv_i8_kernel<true>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a);
    else if (d->dtype == B200_I8)
        launch_kernel(dwconv_i8_kernel<false>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a);
    else
        launch_kernel(dwconv_f16_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
