// pool.cu -- maxpool / avgpool / global avgpool on pixel-major tensors.
// One thread owns 16 bytes of channels of one output pixel and walks the window in (y, x)
// order with one f32 accumulator per channel, which is exactly the reference's sequence
// (source/reference/averagepool.c:71-121, maxpool.c:64-105, global_averagepool.c:21 through
// siso_callback_base utils.c:609: dequantise -> f32 op -> requantise), so int8 results are
// bit-exact.  Adjacent threads own adjacent channel chunks: every load is a coalesced
// 128-bit vector.
#include <float.h>

#include <stdlib.h>

#include "common.cuh"

namespace b200 {

struct PoolArgs {
    int n, c, cp, h, w, oh, ow, kh, kw, sh, sw, pt, pl;
    int is_avg, count_include_pad;
    float s_in, s_out;
    int zp_in, zp_out;
    const void *in;
    void *out;
};

__global__ void __launch_bounds__(128) pool_i8_kernel(const PoolArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    const int chunks = a.cp / 16;
    const long long total = static_cast<long long>(a.n) * a.oh * a.ow * chunks;
    const int8_t *in = static_cast<const int8_t *>(a.in);
    int8_t *out = static_cast<int8_t *>(a.out);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % chunks);
        const int ox = static_cast<int>((i / chunks) % a.ow);
        const int oy = static_cast<int>((i / (static_cast<long long>(chunks) * a.ow)) % a.oh);
        const int b = static_cast<int>(i / (static_cast<long long>(chunks) * a.ow * a.oh));
        const int x0 = ox * a.sw - a.pl, y0 = oy * a.sh - a.pt;
        const int fx0 = max(0, -x0), fx1 = min(a.kw, a.w - x0);
        const int fy0 = max(0, -y0), fy1 = min(a.kh, a.h - y0);
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] = a.is_avg ? 0.f : -FLT_MAX;
        float cnt = 0.f;
        for (int fy = fy0; fy < fy1; fy++) {
            for (int fx = fx0; fx < fx1; fx++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(
                    in + ((static_cast<long long>(b) * a.h + y0 + fy) * a.w + x0 + fx) * a.cp +
                    ch * 16));
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float f =
                            dequant_i8(static_cast<int8_t>(wv[q] >> (8 * e)), a.s_in, a.zp_in);
                        float &t = acc[q * 4 + e];
                        t = a.is_avg ? __fadd_rn(t, f) : fmaxf(t, f);
                    }
                cnt += 1.f;
            }
        }
        if (a.count_include_pad) cnt = static_cast<float>(a.kh * a.kw);
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int j = q * 4 + e;
                const float r = a.is_avg ? __fdiv_rn(acc[j], cnt) : acc[j];
                v[e] = ch * 16 + j < a.c ? quant_i8_exact(r, a.s_out, a.zp_out) : 0;
            }
            pk[q] = pack4_i8(v[0], v[1], v[2], v[3]);
        }
        *reinterpret_cast<uint4 *>(
            out + ((static_cast<long long>(b) * a.oh + oy) * a.ow + ox) * a.cp + ch * 16) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

__global__ void __launch_bounds__(128) pool_f16_kernel(const PoolArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    const int chunks = a.cp / 8;
    const long long total = static_cast<long long>(a.n) * a.oh * a.ow * chunks;
    const __half *in = static_cast<const __half *>(a.in);
    __half *out = static_cast<__half *>(a.out);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % chunks);
        const int ox = static_cast<int>((i / chunks) % a.ow);
        const int oy = static_cast<int>((i / (static_cast<long long>(chunks) * a.ow)) % a.oh);
        const int b = static_cast<int>(i / (static_cast<long long>(chunks) * a.ow * a.oh));
        const int x0 = ox * a.sw - a.pl, y0 = oy * a.sh - a.pt;
        const int fx0 = max(0, -x0), fx1 = min(a.kw, a.w - x0);
        const int fy0 = max(0, -y0), fy1 = min(a.kh, a.h - y0);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = a.is_avg ? 0.f : -FLT_MAX;
        float cnt = 0.f;
        for (int fy = fy0; fy < fy1; fy++) {
            for (int fx = fx0; fx < fx1; fx++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(
                    in + ((static_cast<long long>(b) * a.h + y0 + fy) * a.w + x0 + fx) * a.cp +
                    ch * 8));
                const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float2 f = __half22float2(h[q]);
                    acc[q * 2] = a.is_avg ? acc[q * 2] + f.x : fmaxf(acc[q * 2], f.x);
                    acc[q * 2 + 1] = a.is_avg ? acc[q * 2 + 1] + f.y : fmaxf(acc[q * 2 + 1], f.y);
                }
                cnt += 1.f;
            }
        }
        if (a.count_include_pad) cnt = static_cast<float>(a.kh * a.kw);
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float f0 = a.is_avg ? acc[q * 2] / cnt : acc[q * 2];
            float f1 = a.is_avg ? acc[q * 2 + 1] / cnt : acc[q * 2 + 1];
            f0 = ch * 8 + q * 2 < a.c ? f0 : 0.f;
            f1 = ch * 8 + q * 2 + 1 < a.c ? f1 : 0.f;
            __half2 hv = __floats2half2_rn(f0, f1);
            pk[q] = *reinterpret_cast<uint32_t *>(&hv);
        }
        *reinterpret_cast<uint4 *>(
            out + ((static_cast<long long>(b) * a.oh + oy) * a.ow + ox) * a.cp + ch * 8) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}


// int8 max pooling on bytes: dequantisation is monotone, so max over the dequantised taps is the
// dequantised max of the int8 taps, and dequant -> requant of that one value is a 256-entry table
// (built per CTA with the same device float sequence).  Four channels per __vmaxs4, no float work
// per tap; replaces the float loop above for the max case (ResNet-50 stem pool: 243 us -> see DESIGN).
__global__ void __launch_bounds__(128) maxpool_i8_bytes_kernel(const PoolArgs a)
{
    pdl_launch_dependents();
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_lut[i] = static_cast<uint8_t>(quant_i8_exact(dequant_i8(i - 128, a.s_in, a.zp_in), a.s_out, a.zp_out));
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __syncthreads();
    const int chunks = a.cp / 16;
    const int total = a.n * a.oh * a.ow * chunks;  // < 2^31, host-checked
    const int8_t *in = static_cast<const int8_t *>(a.in);
    int8_t *out = static_cast<int8_t *>(a.out);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ch = i % chunks;
        int rest = i / chunks;
        const int ox = rest % a.ow;
        rest /= a.ow;
        const int oy = rest % a.oh;
        const int b = rest / a.oh;
        const int x0 = ox * a.sw - a.pl, y0 = oy * a.sh - a.pt;
        const int fx0 = max(0, -x0), fx1 = min(a.kw, a.w - x0);
        const int fy0 = max(0, -y0), fy1 = min(a.kh, a.h - y0);
        uint32_t m[4] = {0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u};  // -128 in every lane
        for (int fy = fy0; fy < fy1; fy++) {
            const int8_t *row = in + (static_cast<size_t>(b) * a.h + y0 + fy) * a.w * a.cp + ch * 16;
            for (int fx = fx0; fx < fx1; fx++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + static_cast<size_t>(x0 + fx) * a.cp));
                m[0] = __vmaxs4(m[0], v.x), m[1] = __vmaxs4(m[1], v.y);
                m[2] = __vmaxs4(m[2], v.z), m[3] = __vmaxs4(m[3], v.w);
            }
        }
        const bool empty = fy0 >= fy1 || fx0 >= fx1;  // window entirely in the padding: -FLT_MAX quantises to -128
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t u = m[q] ^ 0x80808080u;  // table index = value + 128
            uint32_t o = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t t = ch * 16 + q * 4 + e < a.c ? (empty ? 0x80u : s_lut[(u >> (8 * e)) & 0xFF]) : 0u;
                o |= t << (8 * e);
            }
            pk[q] = o;
        }
        *reinterpret_cast<uint4 *>(out + ((static_cast<size_t>(b) * a.oh + oy) * a.ow + ox) * a.cp + ch * 16) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// The same for K x K windows that always overlap the image (pad < K, no window entirely in the padding): a padded tap
// is replaced by the nearest tap INSIDE the window (clamped coordinates), which cannot change a maximum, so the K * K
// 128-bit loads are unconditional and issued back to back; one block row per (image, output row), no division but the
// one by the chunk count.  With equal input / output qinfo the requantisation table is the identity and is skipped.
// (ResNet-50's 3x3 stride-2 stem pool, 256 x 64 x 112 x 112: see DESIGN.md.)
template <int K, bool IDENT>
__global__ void __launch_bounds__(256) maxpool_kxk_i8_kernel(const PoolArgs a)
{
    pdl_launch_dependents();
    __shared__ uint8_t s_lut[256];
    if (!IDENT)
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
            s_lut[i] = static_cast<uint8_t>(quant_i8_exact(dequant_i8(i - 128, a.s_in, a.zp_in), a.s_out, a.zp_out));
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __syncthreads();
    const int chunks = a.cp / 16;
    const int per_row = a.ow * chunks;
    const int8_t *in = static_cast<const int8_t *>(a.in);
    int8_t *out = static_cast<int8_t *>(a.out);
    const int rows = a.n * a.oh;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        const int b = row / a.oh, oy = row - b * a.oh;
        const int y0 = oy * a.sh - a.pt;
        int yy[K];
#pragma unroll
        for (int f = 0; f < K; f++) yy[f] = min(max(y0 + f, 0), a.h - 1);
        const int8_t *img = in + static_cast<size_t>(b) * a.h * a.w * a.cp;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
            const int ox = i / chunks, ch = i - ox * chunks;
            const int x0 = ox * a.sw - a.pl;
            uint4 v[K * K];
#pragma unroll
            for (int fy = 0; fy < K; fy++)
#pragma unroll
                for (int fx = 0; fx < K; fx++) {
                    const int xx = min(max(x0 + fx, 0), a.w - 1);
                    v[fy * K + fx] = __ldg(reinterpret_cast<const uint4 *>(img + (static_cast<size_t>(yy[fy]) * a.w + xx) * a.cp + ch * 16));
                }
            uint32_t m[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
            for (int t = 1; t < K * K; t++) {
                m[0] = __vmaxs4(m[0], v[t].x), m[1] = __vmaxs4(m[1], v[t].y);
                m[2] = __vmaxs4(m[2], v[t].z), m[3] = __vmaxs4(m[3], v[t].w);
            }
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t o = m[q];
                if (!IDENT) {
                    const uint32_t u = m[q] ^ 0x80808080u;  // table index = value + 128
                    const uint32_t b0 = s_lut[u & 0xFF], b1 = s_lut[(u >> 8) & 0xFF], b2 = s_lut[(u >> 16) & 0xFF], b3 = s_lut[u >> 24];
                    o = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
                }
                // channels past c (padding lanes of the last chunk) are stored as zeros
                const int left = a.c - (ch * 16 + q * 4);
                if (left < 4) o = left <= 0 ? 0u : (o & (0xFFFFFFFFu >> (8 * (4 - left))));
                pk[q] = o;
            }
            *reinterpret_cast<uint4 *>(out + ((static_cast<size_t>(row)) * a.ow + ox) * a.cp + ch * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// Global average pool, int8: one thread per 32-bit word of channels (four channels), so that a
// 7x7x1024 map spreads over n*256 threads instead of n*64, every load is a coalesced word and the
// h*w loads of a thread are independent (only the four f32 sums are sequential, in (y, x) order
// as averagepool.c:100-109 prescribes).
__global__ void __launch_bounds__(256) gap_i8_kernel(const PoolArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    const int words = a.cp / 4;
    const int total = a.n * words;
    const int hw = a.h * a.w;
    const int8_t *in = static_cast<const int8_t *>(a.in);
    int8_t *out = static_cast<int8_t *>(a.out);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int wd = i % words, b = i / words;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(in + static_cast<size_t>(b) * hw * a.cp) + wd;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // loads in batches of 16 independent words (latency paid once per batch, not per pixel);
        // the sums stay sequential in (y, x) order
        for (int p0 = 0; p0 < hw; p0 += 16) {
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; j++)
                v[j] = p0 + j < hw ? __ldg(src + static_cast<size_t>(p0 + j) * words) : 0u;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (p0 + j < hw) {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        acc[e] = __fadd_rn(acc[e], dequant_i8(static_cast<int8_t>(v[j] >> (8 * e)), a.s_in, a.zp_in));
                }
            }
        }
        const float cnt = static_cast<float>(hw);
        int q[4];
#pragma unroll
        for (int e = 0; e < 4; e++) q[e] = wd * 4 + e < a.c ? quant_i8_exact(__fdiv_rn(acc[e], cnt), a.s_out, a.zp_out) : 0;
        *reinterpret_cast<uint32_t *>(out + static_cast<size_t>(b) * a.cp + wd * 4) = pack4_i8(q[0], q[1], q[2], q[3]);
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_pool2d(const b200_pool_desc *d, void *stream)
{
    if (!d || !d->in || !d->out) {
        set_error("b200_pool2d: null descriptor field");
        return B200_ERR_ARG;
    }
    const int eb = d->dtype == B200_I8 ? 1 : 2;
    if ((d->dtype != B200_I8 && d->dtype != B200_F16) || d->n <= 0 || d->c <= 0 || d->cp < d->c ||
        (d->cp * eb) % 16 || d->kh <= 0 || d->kw <= 0 || d->stride_h <= 0 || d->stride_w <= 0 ||
        d->oh <= 0 || d->ow <= 0) {
        set_error("b200_pool2d: bad descriptor (c=%d cp=%d k=%dx%d)", d->c, d->cp, d->kh, d->kw);
        return B200_ERR_ARG;
    }
    PoolArgs a;
    a.n = d->n, a.c = d->c, a.cp = d->cp, a.h = d->h, a.w = d->w, a.oh = d->oh, a.ow = d->ow;
    a.kh = d->kh, a.kw = d->kw, a.sh = d->stride_h, a.sw = d->stride_w;
    a.pt = d->pad_top, a.pl = d->pad_left;
    a.is_avg = d->is_avg, a.count_include_pad = d->count_include_pad;
    a.s_in = d->s_in, a.s_out = d->s_out, a.zp_in = d->zp_in, a.zp_out = d->zp_out;
    a.in = d->in, a.out = d->out;
    const long long total = static_cast<long long>(d->n) * d->oh * d->ow * (d->cp * eb / 16);
    long long g = (total + 127) / 128;
    const long long cap = static_cast<long long>(sm_count()) * 32;
    const int grid = static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
    if (d->dtype == B200_I8 && d->is_avg && d->oh == 1 && d->ow == 1 && d->kh == d->h && d->kw == d->w &&
        d->pad_top == 0 && d->pad_left == 0) {
        const int tot = d->n * (d->cp / 4);
        launch_kernel(gap_i8_kernel, dim3((tot + 127) / 128), dim3(128), 0, (cudaStream_t)stream, a);
    } else if (d->dtype == B200_I8 && !d->is_avg && d->kh == d->kw && (d->kh == 3 || d->kh == 2) && d->pad_top < d->kh &&
               d->pad_left < d->kw && (d->oh - 1) * d->stride_h - d->pad_top < d->h && (d->ow - 1) * d->stride_w - d->pad_left < d->w &&
               !getenv("SHL_B200_POOL_GENERIC")) {
        // every window overlaps the image: the unrolled kernel with clamped taps
        const bool ident = d->s_in == d->s_out && d->zp_in == d->zp_out;
        const int per_row = d->ow * (d->cp / 16);
        const long long rows = static_cast<long long>(d->n) * d->oh;
        const dim3 g2((per_row + 255) / 256, static_cast<unsigned>(rows < 65535 ? rows : 65535));
        cudaStream_t st = (cudaStream_t)stream;
        if (d->kh == 3) {
            if (ident) launch_kernel(maxpool_kxk_i8_kernel<3, true>, g2, dim3(256), 0, st, a);
            else launch_kernel(maxpool_kxk_i8_kernel<3, false>, g2, dim3(256), 0, st, a);
        } else {
            if (ident) launch_kernel(maxpool_kxk_i8_kernel<2, true>, g2, dim3(256), 0, st, a);
            else launch_kernel(maxpool_kxk_i8_kernel<2, false>, g2, dim3(256), 0, st, a);
        }
    } else if (d->dtype == B200_I8 && !d->is_avg && total < (1ll << 31)) {
        launch_kernel(maxpool_i8_bytes_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a);
    } else if (d->dtype == B200_I8)
        launch_kernel(pool_i8_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a);
    else
        launch_kernel(pool_f16_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
