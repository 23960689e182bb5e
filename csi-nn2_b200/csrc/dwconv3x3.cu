// dwconv3x3.cu -- int8 depthwise 3x3 (stride 1 / 2, dilation 1) on pixel-major tensors: the
// HBM-bound half of a MobileNet block, written so that the instruction stream rather than the
// memory system is what limits it.
//
// One thread owns FOUR channels (one 32-bit word of the channel axis) of a strip of output
// columns, R output rows at a time, so a warp reads / writes 128 contiguous bytes per access
// when C >= 128.  For every input column it loads the S*R+2 input words of that column, turns
// each vertical triple of rows into four "tap words" (x[r][c], x[r+1][c], x[r+2][c], -) with
// byte permutes, and feeds them to dp4a against weight words (w[0][kx][c], w[1][kx][c],
// w[2][kx][c], 0): three dp4a per output instead of nine multiply-adds, no unpacking.  The
// column's contribution is scattered into three rotating sets of accumulators (output columns
// x-1, x, x+1 for stride 1), the completed one is requantised and stored, and the loads of the
// column two steps ahead are already in flight (register ring) to cover HBM latency.
//
// Accumulators start at ibias[c] + kMagicI, so int -> float is one FADD and round-half-even back
// is one FADD + IADD (see common.cuh); the result is the contract of include/b200nn.h bit for bit.
//
// Replaces shl_rvv_dwconv3x3s1_int8 / shl_rvv_dwconv3x3s2_int8
// (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31,244); semantics
// shl_ref_depthwise_conv2d_quant (source/reference/convolution.c:416).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

struct Dw3Args {
    int n, c, cp, h, w, oh, ow, pt, pl;
    int strips, strip_w;  // output columns are cut into `strips` strips of `strip_w`
    const int8_t *in;
    const uint32_t *wcol;  // [3 (kx)][cp] words: (w[0][kx][c], w[1][kx][c], w[2][kx][c], 0)
    int8_t *out;
    int zp_in;
    EpiScalars ep;
};

// tap words of four channels from three vertically adjacent input words
__device__ __forceinline__ void taps_from_rows(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t (&v)[4])
{
    const uint32_t lo = __byte_perm(r0, r1, 0x5140);  // (r0.c0, r1.c0, r0.c1, r1.c1)
    const uint32_t hi = __byte_perm(r0, r1, 0x7362);  // (r0.c2, r1.c2, r0.c3, r1.c3)
    v[0] = __byte_perm(lo, r2, 0x4410);
    v[1] = __byte_perm(lo, r2, 0x5532);
    v[2] = __byte_perm(hi, r2, 0x6610);
    v[3] = __byte_perm(hi, r2, 0x7732);
}

template <int S, int R>
struct Dw3 {
    static constexpr int kRows = S * (R - 1) + 3;  // input rows feeding R output rows

    const Dw3Args &a;
    int b, oy0, cw;         // image, first output row, channel word
    uint32_t wk[3][4];      // weight words per kx, per channel of the word
    float mu[4], ba[4];
    int init[4];            // ibias + kMagicI
    int zp_m;               // zp_out - kMagicI
    const uint8_t *lut;
    uint32_t padw;

    __device__ __forceinline__ Dw3(const Dw3Args &args) : a(args) {}

    __device__ __forceinline__ void load_col(int xi, uint32_t (&rows)[kRows]) const
    {
        const int iy0 = oy0 * S - a.pt;
        const bool xok = xi >= 0 && xi < a.w;
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            const int iy = iy0 + r;
            uint32_t v = padw;
            if (xok && iy >= 0 && iy < a.h)
                v = __ldg(reinterpret_cast<const uint32_t *>(
                    a.in + ((static_cast<long long>(b) * a.h + iy) * a.w + xi) * a.cp + cw * 4));
            rows[r] = v;
        }
    }

    __device__ __forceinline__ void reset(int (&acc)[R][4]) const
    {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[r][e] = init[e];
    }

    // add this input column's contribution with kernel column kx
    __device__ __forceinline__ void mac(int (&acc)[R][4], const uint32_t (&v)[R][4], int kx) const
    {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int e = 0; e < 4; e++)
                acc[r][e] = __dp4a(static_cast<int>(v[r][e]), static_cast<int>(wk[kx][e]), acc[r][e]);
    }

    __device__ __forceinline__ void finish(int (&acc)[R][4], int ox, int x_lo, int x_hi) const
    {
        if (ox >= x_lo && ox < x_hi) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int oy = oy0 + r;
                if (oy >= a.oh) break;
                int q[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float f = fmaf(magic_to_float(acc[r][e]), mu[e], ba[e]);
                    q[e] = magic_round(f, zp_m);
                    if (a.ep.act != B200_ACT_NONE) q[e] = max(q[e], a.ep.zp_out);
                    if (a.ep.act == B200_ACT_RELU6) q[e] = min(q[e], a.ep.q6);
                }
                const uint32_t word = lut ? lut4_i8(q[0], q[1], q[2], q[3], lut)
                                          : pack4_sat_i8(q[0], q[1], q[2], q[3]);
                *reinterpret_cast<uint32_t *>(
                    a.out + ((static_cast<long long>(b) * a.oh + oy) * a.ow + ox) * a.cp + cw * 4) = word;
            }
        }
        reset(acc);
    }

    __device__ __forceinline__ void column_taps(const uint32_t (&rows)[kRows], uint32_t (&v)[R][4]) const
    {
#pragma unroll
        for (int r = 0; r < R; r++) taps_from_rows(rows[r * S], rows[r * S + 1], rows[r * S + 2], v[r]);
    }

    // One input column xi.  accP / accC / accN are the accumulators of the previous, current
    // and next output column relative to this input column (see run()).
    __device__ __forceinline__ void step(int xi, const uint32_t (&rows)[kRows], int (&accP)[R][4],
                                         int (&accC)[R][4], int (&accN)[R][4], int x_lo, int x_hi) const
    {
        uint32_t v[R][4];
        column_taps(rows, v);
        if (S == 1) {
            // input column xi (= ox - pl + kx): outputs ox = xi+pl-kx for kx = 0, 1, 2
            const int oc = xi + a.pl;  // output column that sees this input under kx = 0
            mac(accN, v, 0);           // output oc
            mac(accC, v, 1);           // output oc - 1
            mac(accP, v, 2);           // output oc - 2: complete now
            finish(accP, oc - 2, x_lo, x_hi);
        }
    }
};

// stride 1: R rows x strip of columns per thread, register ring of 3 columns
template <int R>
__global__ void __launch_bounds__(R == 4 ? 128 : 256, R == 4 ? 3 : 2) dw3x3s1_i8_kernel(const Dw3Args a)
{
    __shared__ uint8_t s_lut[256];
    if (a.ep.post_lut != nullptr)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = static_cast<uint8_t>(a.ep.post_lut[i]);
    __syncthreads();

    using K = Dw3<1, R>;
    const int words = a.cp / 4;
    const int rgroups = (a.oh + R - 1) / R;
    const long long total = static_cast<long long>(a.n) * rgroups * a.strips * words;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        K k(a);
        k.cw = static_cast<int>(i % words);
        const int strip = static_cast<int>((i / words) % a.strips);
        const int rg = static_cast<int>((i / (static_cast<long long>(words) * a.strips)) % rgroups);
        k.b = static_cast<int>(i / (static_cast<long long>(words) * a.strips * rgroups));
        k.oy0 = rg * R;
        k.lut = a.ep.post_lut != nullptr ? s_lut : nullptr;
        k.padw = 0x01010101u * static_cast<uint32_t>(a.zp_in & 0xFF);
        k.zp_m = a.ep.zp_out - kMagicI;
        {
            const uint4 w0 = __ldg(reinterpret_cast<const uint4 *>(a.wcol + 0 * a.cp + k.cw * 4));
            const uint4 w1 = __ldg(reinterpret_cast<const uint4 *>(a.wcol + 1 * a.cp + k.cw * 4));
            const uint4 w2 = __ldg(reinterpret_cast<const uint4 *>(a.wcol + 2 * a.cp + k.cw * 4));
            k.wk[0][0] = w0.x, k.wk[0][1] = w0.y, k.wk[0][2] = w0.z, k.wk[0][3] = w0.w;
            k.wk[1][0] = w1.x, k.wk[1][1] = w1.y, k.wk[1][2] = w1.z, k.wk[1][3] = w1.w;
            k.wk[2][0] = w2.x, k.wk[2][1] = w2.y, k.wk[2][2] = w2.z, k.wk[2][3] = w2.w;
            const float4 m4 = __ldg(reinterpret_cast<const float4 *>(a.ep.mult + k.cw * 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.ep.badd + k.cw * 4));
            const int4 i4 = __ldg(reinterpret_cast<const int4 *>(a.ep.ibias + k.cw * 4));
            k.mu[0] = m4.x, k.mu[1] = m4.y, k.mu[2] = m4.z, k.mu[3] = m4.w;
            k.ba[0] = b4.x, k.ba[1] = b4.y, k.ba[2] = b4.z, k.ba[3] = b4.w;
            k.init[0] = i4.x + kMagicI, k.init[1] = i4.y + kMagicI, k.init[2] = i4.z + kMagicI,
            k.init[3] = i4.w + kMagicI;
        }
        const int x_lo = strip * a.strip_w, x_hi = min(a.ow, x_lo + a.strip_w);
        // input columns x_lo - pl ... x_hi - 1 - pl + 2
        const int xi0 = x_lo - a.pl, xi1 = x_hi - a.pl + 2;  // [xi0, xi1)
        int accA[R][4], accB[R][4], accC[R][4];
        k.reset(accA), k.reset(accB), k.reset(accC);
        uint32_t buf0[K::kRows], buf1[K::kRows], buf2[K::kRows];
        k.load_col(xi0, buf0);
        k.load_col(xi0 + 1, buf1);
        for (int xi = xi0; xi < xi1; xi += 3) {
            k.load_col(xi + 2, buf2);
            k.step(xi, buf0, accA, accB, accC, x_lo, x_hi);
            if (xi + 1 >= xi1) break;
            k.load_col(xi + 3, buf0);
            k.step(xi + 1, buf1, accB, accC, accA, x_lo, x_hi);
            if (xi + 2 >= xi1) break;
            k.load_col(xi + 4, buf1);
            k.step(xi + 2, buf2, accC, accA, accB, x_lo, x_hi);
        }
    }
}

// stride 2: output column ox reads input columns 2*ox - pl + {0, 1, 2}
template <int R>
__global__ void __launch_bounds__(256, 2) dw3x3s2_i8_kernel(const Dw3Args a)
{
    __shared__ uint8_t s_lut[256];
    if (a.ep.post_lut != nullptr)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = static_cast<uint8_t>(a.ep.post_lut[i]);
    __syncthreads();

    using K = Dw3<2, R>;
    const int words = a.cp / 4;
    const int rgroups = (a.oh + R - 1) / R;
    const long long total = static_cast<long long>(a.n) * rgroups * a.strips * words;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        K k(a);
        k.cw = static_cast<int>(i % words);
        const int strip = static_cast<int>((i / words) % a.strips);
        const int rg = static_cast<int>((i / (static_cast<long long>(words) * a.strips)) % rgroups);
        k.b = static_cast<int>(i / (static_cast<long long>(words) * a.strips * rgroups));
        k.oy0 = rg * R;
        k.lut = a.ep.post_lut != nullptr ? s_lut : nullptr;
        k.padw = 0x01010101u * static_cast<uint32_t>(a.zp_in & 0xFF);
        k.zp_m = a.ep.zp_out - kMagicI;
        {
            const uint4 w0 = __ldg(reinterpret_cast<const uint4 *>(a.wcol + 0 * a.cp + k.cw * 4));
            const uint4 w1 = __ldg(reinterpret_cast<const uint4 *>(a.wcol + 1 * a.cp + k.cw * 4));
            const uint4 w2 = __ldg(reinterpret_cast<const uint4 *>(a.wcol + 2 * a.cp + k.cw * 4));
            k.wk[0][0] = w0.x, k.wk[0][1] = w0.y, k.wk[0][2] = w0.z, k.wk[0][3] = w0.w;
            k.wk[1][0] = w1.x, k.wk[1][1] = w1.y, k.wk[1][2] = w1.z, k.wk[1][3] = w1.w;
            k.wk[2][0] = w2.x, k.wk[2][1] = w2.y, k.wk[2][2] = w2.z, k.wk[2][3] = w2.w;
            const float4 m4 = __ldg(reinterpret_cast<const float4 *>(a.ep.mult + k.cw * 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.ep.badd + k.cw * 4));
            const int4 i4 = __ldg(reinterpret_cast<const int4 *>(a.ep.ibias + k.cw * 4));
            k.mu[0] = m4.x, k.mu[1] = m4.y, k.mu[2] = m4.z, k.mu[3] = m4.w;
            k.ba[0] = b4.x, k.ba[1] = b4.y, k.ba[2] = b4.z, k.ba[3] = b4.w;
            k.init[0] = i4.x + kMagicI, k.init[1] = i4.y + kMagicI, k.init[2] = i4.z + kMagicI,
            k.init[3] = i4.w + kMagicI;
        }
        const int x_lo = strip * a.strip_w, x_hi = min(a.ow, x_lo + a.strip_w);
        int acc[R][4];
        k.reset(acc);
        // three input columns per output column; the last one of ox is the first one of ox + 1
        uint32_t c0[K::kRows], c1[K::kRows], c2[K::kRows], n1[K::kRows], n2[K::kRows];
        uint32_t v[R][4];
        int xi = 2 * x_lo - a.pl;
        k.load_col(xi, c0);
        k.load_col(xi + 1, c1);
        k.load_col(xi + 2, c2);
        for (int ox = x_lo; ox < x_hi; ox++, xi += 2) {
            // prefetch the two new columns of the next output while this one is computed
            k.load_col(xi + 3, n1);
            k.load_col(xi + 4, n2);
            k.column_taps(c0, v);
            k.mac(acc, v, 0);
            k.column_taps(c1, v);
            k.mac(acc, v, 1);
            k.column_taps(c2, v);
            k.mac(acc, v, 2);
            k.finish(acc, ox, x_lo, x_hi);
#pragma unroll
            for (int r = 0; r < K::kRows; r++) c0[r] = c2[r], c1[r] = n1[r], c2[r] = n2[r];
        }
    }
}

}  // namespace b200

using namespace b200;

// called by b200_dwconv2d (dwconv.cu) for the int8 3x3 shapes it covers; `wcol` is the
// kx-major repack of the depthwise weights (b200_opt/quant.c builds it next to the generic one)
int b200_dwconv3x3_i8_launch(const b200_dwconv_desc *d, const void *wcol, void *stream)
{
    Dw3Args a;
    a.n = d->n, a.c = d->c, a.cp = d->cp, a.h = d->h, a.w = d->w, a.oh = d->oh, a.ow = d->ow;
    a.pt = d->pad_top, a.pl = d->pad_left;
    a.in = static_cast<const int8_t *>(d->in);
    a.wcol = static_cast<const uint32_t *>(wcol);
    a.out = static_cast<int8_t *>(d->out);
    a.zp_in = d->zp_in;
    a.ep = make_epi(d->ep);
    const int words = d->cp / 4;
    const int stride = d->stride_h;
    // rows per thread (stride 1): 3 keeps the three accumulator sets, the column ring and the tap
    // words inside 128 registers without spilling (4 spills); SHL_B200_DW_ROWS overrides for tuning
    const char *e = getenv("SHL_B200_DW_ROWS");
    int rows_s1 = e ? atoi(e) : 3;
    if (rows_s1 < 2 || rows_s1 > 4) rows_s1 = 3;
    const int R = stride == 1 ? rows_s1 : 2;
    const int rgroups = (d->oh + R - 1) / R;
    // enough independent strips to fill the machine (2 CTAs x 256 threads per SM), but strips as
    // long as possible: each strip re-reads two halo columns (stride 1)
    const long long want = static_cast<long long>(sm_count()) * 512 * 2;
    long long base = static_cast<long long>(d->n) * rgroups * words;
    int strips = 1;
    while (base * strips < want && d->ow / (strips + 1) >= 7) strips++;
    a.strips = strips;
    a.strip_w = (d->ow + strips - 1) / strips;
    a.strips = (d->ow + a.strip_w - 1) / a.strip_w;
    const long long total = base * a.strips;
    const int threads = (stride == 1 && R == 4) ? 128 : 256;
    long long g = (total + threads - 1) / threads;
    const long long cap = static_cast<long long>(sm_count()) * 2 * 8 * (256 / threads);
    const int grid = static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
    if (stride == 1 && R == 4)
        dw3x3s1_i8_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else if (stride == 1 && R == 3)
        dw3x3s1_i8_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else if (stride == 1)
        dw3x3s1_i8_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else
        dw3x3s2_i8_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
