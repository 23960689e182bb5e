// layout.cu -- layout conversion at the API boundary and im2col.
//   b200_nchw_to_nhwc / b200_nhwc_to_nchw : the API's NCHW tensors <-> the device's pixel-major
//     [N][H][W][Cp] (stand-ins for the NCHW <-> NC1HWC0 reorders of source/thead_rvv/data_convert.c)
//   b200_im2col : the gather half of "im2col + GEMM conv2d" (k order = (ky, kx, ci), ci fastest,
//     so that one tap is a contiguous run of channels in the pixel-major input).
// All three are pure data movement: coalesced along the contiguous axis of whichever side is
// wider, 16-byte vectors on the pixel-major side.
#include "common.cuh"

namespace b200 {

// one thread: 16 bytes of channels of one pixel.  Adjacent threads = adjacent pixels, so the
// strided NCHW side is coalesced across the warp (one byte / half per lane per channel).
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T *__restrict__ src, T *__restrict__ dst, int n, int c,
                                    int hw, int cp, T pad)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    constexpr int V = 16 / sizeof(T);
    const int chunks = cp / V;
    const long long total = static_cast<long long>(n) * chunks * hw;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i % hw);
        const int ch = static_cast<int>((i / hw) % chunks);
        const int b = static_cast<int>(i / (static_cast<long long>(hw) * chunks));
        alignas(16) T v[V];
#pragma unroll
        for (int j = 0; j < V; j++) {
            const int cc = ch * V + j;
            v[j] = cc < c ? src[(static_cast<long long>(b) * c + cc) * hw + p] : pad;
        }
        *reinterpret_cast<uint4 *>(dst + (static_cast<long long>(b) * hw + p) * cp + ch * V) =
            *reinterpret_cast<const uint4 *>(v);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ src, T *__restrict__ dst, int n, int c,
                                    int hw, int cp)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    constexpr int V = 16 / sizeof(T);
    const int chunks = cp / V;
    const long long total = static_cast<long long>(n) * chunks * hw;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i % hw);
        const int ch = static_cast<int>((i / hw) % chunks);
        const int b = static_cast<int>(i / (static_cast<long long>(hw) * chunks));
        alignas(16) T v[V];
        *reinterpret_cast<uint4 *>(v) = *reinterpret_cast<const uint4 *>(
            src + (static_cast<long long>(b) * hw + p) * cp + ch * V);
#pragma unroll
        for (int j = 0; j < V; j++) {
            const int cc = ch * V + j;
            if (cc < c) dst[(static_cast<long long>(b) * c + cc) * hw + p] = v[j];
        }
    }
}

struct Im2colArgs {
    int n, h, w, cp_in, in_nchw, c_total, c_off, cg;
    int oh, ow, kh, kw, sh, sw, pt, pl, dh, dw;
    int ldk;
    const void *in;
    void *col;
};

// one thread: one 16-byte vector of one im2col row.
template <typename T>
__global__ void im2col_kernel(const Im2colArgs a, T pad)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    constexpr int V = 16 / sizeof(T);
    const int vecs = a.ldk / V;
    const int kvalid = a.kh * a.kw * a.cg;
    const long long rows = static_cast<long long>(a.n) * a.oh * a.ow;
    const long long total = rows * vecs;
    const T *in = static_cast<const T *>(a.in);
    T *col = static_cast<T *>(a.col);
    const bool fast = !a.in_nchw && (a.cg % V == 0) && (a.c_off % V == 0);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int vec = static_cast<int>(i % vecs);
        const long long m = i / vecs;
        const int ox = static_cast<int>(m % a.ow);
        const int oy = static_cast<int>((m / a.ow) % a.oh);
        const int b = static_cast<int>(m / (static_cast<long long>(a.ow) * a.oh));
        alignas(16) T v[V];
        const int k0 = vec * V;
        if (fast) {
            // the vector lies inside one tap: a contiguous run of channels
            if (k0 < kvalid) {
                const int tap = k0 / a.cg, ci = k0 % a.cg;
                const int iy = oy * a.sh - a.pt + (tap / a.kw) * a.dh;
                const int ix = ox * a.sw - a.pl + (tap % a.kw) * a.dw;
                if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) {
                    *reinterpret_cast<uint4 *>(v) = *reinterpret_cast<const uint4 *>(
                        in + ((static_cast<long long>(b) * a.h + iy) * a.w + ix) * a.cp_in + a.c_off +
                        ci);
                } else {
#pragma unroll
                    for (int j = 0; j < V; j++) v[j] = pad;
                }
            } else {
#pragma unroll
                for (int j = 0; j < V; j++) v[j] = pad;
            }
        } else {
#pragma unroll
            for (int j = 0; j < V; j++) {
                const int k = k0 + j;
                T x = pad;
                if (k < kvalid) {
                    const int tap = k / a.cg, ci = k % a.cg;
                    const int iy = oy * a.sh - a.pt + (tap / a.kw) * a.dh;
                    const int ix = ox * a.sw - a.pl + (tap % a.kw) * a.dw;
                    if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) {
                        x = a.in_nchw
                                ? in[((static_cast<long long>(b) * a.c_total + a.c_off + ci) * a.h + iy) *
                                         a.w + ix]
                                : in[((static_cast<long long>(b) * a.h + iy) * a.w + ix) * a.cp_in +
                                     a.c_off + ci];
                    }
                }
                v[j] = x;
            }
        }
        *reinterpret_cast<uint4 *>(col + m * a.ldk + k0) = *reinterpret_cast<const uint4 *>(v);
    }
}

static int grid_for(long long total, int block)
{
    long long g = (total + block - 1) / block;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b200

using namespace b200;

extern "C" int b200_nchw_to_nhwc(const void *src, void *dst, int n, int c, int h, int w, int cp,
                                 int elem_bytes, int pad, void *stream)
{
    if (!src || !dst || n <= 0 || c <= 0 || h <= 0 || w <= 0 || cp < c ||
        (elem_bytes != 1 && elem_bytes != 2) || (cp * elem_bytes) % 16) {
        set_error("b200_nchw_to_nhwc: bad arguments (n=%d c=%d h=%d w=%d cp=%d elem=%d)", n, c, h, w,
                  cp, elem_bytes);
        return B200_ERR_ARG;
    }
    const long long total = static_cast<long long>(n) * h * w * (cp * elem_bytes / 16);
    const int grid = grid_for(total, 256);
    if (elem_bytes == 1)
        launch_kernel(nchw_to_nhwc_kernel<int8_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const int8_t *>(src), static_cast<int8_t *>(dst), n, c, h * w, cp,
            static_cast<int8_t>(pad));
    else
        launch_kernel(nchw_to_nhwc_kernel<uint16_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), n, c, h * w, cp,
            static_cast<uint16_t>(pad));
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_nhwc_to_nchw(const void *src, void *dst, int n, int c, int h, int w, int cp,
                                 int elem_bytes, void *stream)
{
    if (!src || !dst || n <= 0 || c <= 0 || h <= 0 || w <= 0 || cp < c ||
        (elem_bytes != 1 && elem_bytes != 2) || (cp * elem_bytes) % 16) {
        set_error("b200_nhwc_to_nchw: bad arguments (n=%d c=%d h=%d w=%d cp=%d elem=%d)", n, c, h, w,
                  cp, elem_bytes);
        return B200_ERR_ARG;
    }
    const long long total = static_cast<long long>(n) * h * w * (cp * elem_bytes / 16);
    const int grid = grid_for(total, 256);
    if (elem_bytes == 1)
        launch_kernel(nhwc_to_nchw_kernel<int8_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const int8_t *>(src), static_cast<int8_t *>(dst), n, c, h * w, cp);
    else
        launch_kernel(nhwc_to_nchw_kernel<uint16_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
            static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), n, c, h * w, cp);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

namespace b200 {

// Pixel-major input with whole 16-byte vectors per tap (the common case: every conv but a network's
// first layer).  A thread keeps ONE vector position of the K row -- its tap and channel offset are
// computed once -- and walks the output pixels with a fixed stride, carrying (image, y, x) along as
// digits: no division per vector (the generic kernel below spends ~7 integer divisions, 64-bit, per
// 16 bytes, which made this gather run at a tenth of HBM speed).
template <typename T>
__global__ void __launch_bounds__(256) im2col_walk_kernel(const Im2colArgs a, T pad, int rows_per_block, int drow_b,
                                                          int drow_y, int drow_x)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    constexpr int V = 16 / sizeof(T);
    const int vecs = a.ldk / V;
    const int v = threadIdx.x % vecs, r = threadIdx.x / vecs;
    if (r >= rows_per_block) return;
    const int k0 = v * V;
    const bool kvalid = k0 < a.kh * a.kw * a.cg;
    const int tap = kvalid ? k0 / a.cg : 0, ci = kvalid ? k0 % a.cg : 0;
    const int dy = (tap / a.kw) * a.dh - a.pt, dx = (tap % a.kw) * a.dw - a.pl;
    const T *in = static_cast<const T *>(a.in) + a.c_off + ci;
    T *col = static_cast<T *>(a.col) + k0;
    const long long rows = static_cast<long long>(a.n) * a.oh * a.ow;
    long long m = static_cast<long long>(blockIdx.x) * rows_per_block + r;
    if (m >= rows) return;
    int ox = static_cast<int>(m % a.ow);
    int oy = static_cast<int>((m / a.ow) % a.oh);
    int b = static_cast<int>(m / (static_cast<long long>(a.ow) * a.oh));
    uint4 padv;
    {
        alignas(16) T pv[V];
#pragma unroll
        for (int j = 0; j < V; j++) pv[j] = pad;
        padv = *reinterpret_cast<const uint4 *>(pv);
    }
    const long long mstep = static_cast<long long>(gridDim.x) * rows_per_block;
    for (; m < rows; m += mstep) {
        const int iy = oy * a.sh + dy, ix = ox * a.sw + dx;
        uint4 val = padv;
        if (kvalid && iy >= 0 && iy < a.h && ix >= 0 && ix < a.w)
            val = __ldg(reinterpret_cast<const uint4 *>(in + ((static_cast<long long>(b) * a.h + iy) * a.w + ix) * a.cp_in));
        *reinterpret_cast<uint4 *>(col + m * a.ldk) = val;
        // advance (b, oy, ox) by the grid's row stride, digit by digit
        ox += drow_x;
        if (ox >= a.ow) ox -= a.ow, oy++;
        oy += drow_y;
        if (oy >= a.oh) oy -= a.oh, b++;
        b += drow_b;
    }
}

}  // namespace b200

extern "C" int b200_im2col(const b200_im2col_desc *d, void *stream)
{
    if (!d || !d->in || !d->col) {
        set_error("b200_im2col: null descriptor field");
        return B200_ERR_ARG;
    }
    const int eb = d->dtype == B200_I8 ? 1 : 2;
    if ((d->dtype != B200_I8 && d->dtype != B200_F16) || d->n <= 0 || d->cg <= 0 || d->kh <= 0 ||
        d->kw <= 0 || d->oh <= 0 || d->ow <= 0 || d->stride_h <= 0 || d->stride_w <= 0 ||
        d->dil_h <= 0 || d->dil_w <= 0 || (d->ldk * eb) % 16 || d->ldk < d->kh * d->kw * d->cg ||
        (!d->in_nchw && (d->cp_in * eb) % 16)) {
        set_error("b200_im2col: bad descriptor (cg=%d k=%dx%d ldk=%d cp_in=%d)", d->cg, d->kh, d->kw,
                  d->ldk, d->cp_in);
        return B200_ERR_ARG;
    }
    Im2colArgs a;
    a.n = d->n, a.h = d->h, a.w = d->w, a.cp_in = d->cp_in, a.in_nchw = d->in_nchw;
    a.c_total = d->c_total, a.c_off = d->c_off, a.cg = d->cg;
    a.oh = d->oh, a.ow = d->ow, a.kh = d->kh, a.kw = d->kw;
    a.sh = d->stride_h, a.sw = d->stride_w, a.pt = d->pad_top, a.pl = d->pad_left;
    a.dh = d->dil_h, a.dw = d->dil_w, a.ldk = d->ldk, a.in = d->in, a.col = d->col;
    const long long total = static_cast<long long>(d->n) * d->oh * d->ow * (d->ldk * eb / 16);
    const int vecs = d->ldk * eb / 16, V = 16 / eb;
    if (!d->in_nchw && d->cg % V == 0 && d->c_off % V == 0 && vecs <= 256) {
        const int rpb = 256 / vecs;  // output pixels per block
        const long long rows = static_cast<long long>(d->n) * d->oh * d->ow;
        long long g = (rows + rpb - 1) / rpb;
        const long long cap = static_cast<long long>(sm_count()) * 16;
        const int grid = static_cast<int>(g < cap ? g : cap);
        long long step = static_cast<long long>(grid) * rpb;  // rows skipped per iteration, as (image, y, x) digits
        const int dx = static_cast<int>(step % d->ow);
        step /= d->ow;
        const int dy = static_cast<int>(step % d->oh);
        const int db = static_cast<int>(step / d->oh);
        if (eb == 1)
            launch_kernel(im2col_walk_kernel<int8_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a,
                          static_cast<int8_t>(d->pad_value), rpb, db, dy, dx);
        else
            launch_kernel(im2col_walk_kernel<uint16_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a,
                          static_cast<uint16_t>(0), rpb, db, dy, dx);
        B200_LAUNCH_CHECK();
        return B200_OK;
    }
    const int grid = grid_for(total, 256);
    if (eb == 1)
        launch_kernel(im2col_kernel<int8_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a, static_cast<int8_t>(d->pad_value));
    else
        launch_kernel(im2col_kernel<uint16_t>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a, static_cast<uint16_t>(0));
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// ---- concat: one input of a concatenation copied into its slice of the output --------------------
// source/reference/concat.c:20-72: every input is dequantised with its own qinfo, the pieces are
// laid side by side along `axis`, the whole output is quantised with the output qinfo.  For int8
// that is a 256-entry table per input (requant(dequant(q))); fp16 -> f32 -> fp16 is the identity.
// One launch moves one input; V bytes per thread (16 when the channel counts allow it).  On the
// channel axis the last input also zeroes the output's padding lanes.
namespace b200 {

struct ConcatArgs {
    const uint8_t *in;
    uint8_t *out;
    const int8_t *lut;  // NULL: plain copy
    int n, h, w;        // input pixels
    int oh, ow;         // output rows / columns per image
    int n_off, h_off, w_off;
    int in_pitch, out_pitch;  // bytes per pixel
    int row_bytes;            // bytes copied per pixel (input channels)
    int out_byte_off;         // byte offset of the slice inside an output pixel
    int zero_bytes;           // bytes zeroed after the slice (padding lanes)
    int in_byte_off;          // extract: byte offset of the slice inside an input pixel
    int extract;              // 1: the input is the large tensor, the output the slice (split)
    int spatial;              // 1: a pixel offset has to be applied
};

template <typename VT>
__device__ __forceinline__ VT lut_apply(VT v, const uint8_t *s_lut);
template <>
__device__ __forceinline__ uint8_t lut_apply<uint8_t>(uint8_t v, const uint8_t *s_lut)
{
    return s_lut[v ^ 0x80];
}
template <>
__device__ __forceinline__ uint32_t lut_apply<uint32_t>(uint32_t v, const uint8_t *s_lut)
{
    uint32_t o = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) o |= static_cast<uint32_t>(s_lut[((v >> (8 * e)) & 0xFF) ^ 0x80]) << (8 * e);
    return o;
}
template <>
__device__ __forceinline__ uint4 lut_apply<uint4>(uint4 v, const uint8_t *s_lut)
{
    return make_uint4(lut_apply<uint32_t>(v.x, s_lut), lut_apply<uint32_t>(v.y, s_lut),
                      lut_apply<uint32_t>(v.z, s_lut), lut_apply<uint32_t>(v.w, s_lut));
}
template <typename VT>
__device__ __forceinline__ VT vec_zero();
template <>
__device__ __forceinline__ uint8_t vec_zero<uint8_t>() { return 0; }
template <>
__device__ __forceinline__ uint16_t vec_zero<uint16_t>() { return 0; }
template <>
__device__ __forceinline__ uint32_t vec_zero<uint32_t>() { return 0; }
template <>
__device__ __forceinline__ uint4 vec_zero<uint4>() { return make_uint4(0, 0, 0, 0); }

template <typename VT, bool LUT>
__global__ void __launch_bounds__(256) concat_slice_kernel(ConcatArgs a)
{
    pdl_launch_dependents();
    pdl_wait();  // inputs and the output buffer belong to the predecessor until here
    __shared__ uint8_t s_lut[256];
    if constexpr (LUT) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = static_cast<uint8_t>(a.lut[i]);
        __syncthreads();
    }
    constexpr int V = sizeof(VT);
    const int copy_units = a.row_bytes / V;
    const int units = copy_units + a.zero_bytes / V;
    const long long total = static_cast<long long>(a.n) * a.h * a.w * units;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long p = i / units;
        const int u = static_cast<int>(i - p * units);
        long long po = p;
        if (a.spatial) {
            const int x = static_cast<int>(p % a.w);
            const long long r = p / a.w;
            const int y = static_cast<int>(r % a.h);
            const int b = static_cast<int>(r / a.h);
            po = (static_cast<long long>(b + a.n_off) * a.oh + (y + a.h_off)) * a.ow + (x + a.w_off);
        }
        const long long ps = a.extract ? po : p, pd = a.extract ? p : po;  // source / destination pixel
        VT v;
        if (u < copy_units) {
            v = *reinterpret_cast<const VT *>(a.in + ps * a.in_pitch + a.in_byte_off + static_cast<long long>(u) * V);
            if constexpr (LUT) v = lut_apply<VT>(v, s_lut);
        } else {
            v = vec_zero<VT>();
        }
        *reinterpret_cast<VT *>(a.out + pd * a.out_pitch + a.out_byte_off + static_cast<long long>(u) * V) = v;
    }
}

template <typename VT>
static void launch_concat(const ConcatArgs &a, int grid, cudaStream_t stream)
{
    if constexpr (sizeof(VT) != 2) {
        if (a.lut) {
            launch_kernel(concat_slice_kernel<VT, true>, dim3(grid), dim3(256), 0, stream, a);
            return;
        }
    }
    {
        launch_kernel(concat_slice_kernel<VT, false>, dim3(grid), dim3(256), 0, stream, a);
    }
}

}  // namespace b200

extern "C" int b200_concat_slice(const b200_concat_desc *d, void *stream)
{
    using namespace b200;
    const int eb = !d ? 0 : (d->dtype == B200_I8 ? 1 : (d->dtype == B200_F16 ? 2 : 0));
    if (!d || !eb || !d->in || !d->out || d->n <= 0 || d->h <= 0 || d->w <= 0 || d->c <= 0 || d->axis < 0 ||
        d->axis > 3 || d->offset < 0 || (d->extract ? d->cp_out : d->cp_in) < d->c || (d->cp_in * eb) % 16 || (d->cp_out * eb) % 16 ||
        (d->dtype != B200_I8 && d->lut)) {
        set_error("b200_concat_slice: bad descriptor");
        return B200_ERR_ARG;
    }
    const int span[4] = {d->n, d->c, d->h, d->w}, room[4] = {d->on, d->oc, d->oh, d->ow};
    for (int ax = 0; ax < 4; ax++) {
        const int off = ax == d->axis ? d->offset : 0;
        if (off + span[ax] > room[ax] || (ax != d->axis && span[ax] != room[ax])) {
            set_error("b200_concat_slice: slice does not fit the output (axis %d: %d + %d > %d)", ax, off, span[ax],
                      room[ax]);
            return B200_ERR_ARG;
        }
    }
    if ((d->extract ? d->cp_in : d->cp_out) < d->oc) {
        set_error("b200_concat_slice: channel pitch %d of the whole tensor below its channel count %d",
                  d->extract ? d->cp_in : d->cp_out, d->oc);
        return B200_ERR_ARG;
    }
    ConcatArgs a;
    a.in = static_cast<const uint8_t *>(d->in), a.out = static_cast<uint8_t *>(d->out), a.lut = d->lut;
    a.n = d->n, a.h = d->h, a.w = d->w, a.oh = d->oh, a.ow = d->ow;
    a.n_off = d->axis == 0 ? d->offset : 0, a.h_off = d->axis == 2 ? d->offset : 0, a.w_off = d->axis == 3 ? d->offset : 0;
    a.spatial = d->axis != 1;
    a.in_pitch = d->cp_in * eb, a.out_pitch = d->cp_out * eb;
    a.row_bytes = d->c * eb;
    a.extract = d->extract ? 1 : 0;
    a.out_byte_off = (d->axis == 1 && !a.extract) ? d->offset * eb : 0;
    a.in_byte_off = (d->axis == 1 && a.extract) ? d->offset * eb : 0;
    // padding lanes: after the last channel slice, or after every pixel's channels on the other axes; an
    // extracted slice is a whole tensor: always
    const int written_to = a.out_byte_off + a.row_bytes;
    const int is_tail = a.extract || d->axis != 1 || d->offset + d->c == d->oc;
    a.zero_bytes = is_tail ? a.out_pitch - written_to : 0;
    int v = 16;
    while (v > eb && (a.row_bytes % v || a.out_byte_off % v || a.in_byte_off % v || a.zero_bytes % v)) v = v == 16 ? 4 : eb;
    const long long total = static_cast<long long>(d->n) * d->h * d->w * ((a.row_bytes + a.zero_bytes) / v);
    const int grid = grid_for(total, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (v == 16)
        launch_concat<uint4>(a, grid, s);
    else if (v == 4)
        launch_concat<uint32_t>(a, grid, s);
    else if (v == 2)
        launch_concat<uint16_t>(a, grid, s);
    else
        launch_concat<uint8_t>(a, grid, s);
    B200_LAUNCH_CHECK();
    return B200_OK;
}


// ---- row sums for asymmetric weights -----------------------------------------------------------------------
// rs[m] = sum_{k < K} a[m][k] - zp_in * K: one warp per row, 16 bytes per lane and step, dp4a against ones
namespace b200 {
__global__ void __launch_bounds__(256) rowsum_i8_kernel(const int8_t *__restrict__ a, int lda, int m, int k, int zp_in,
                                                        int32_t *__restrict__ rs)
{
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int k16 = k & ~15;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < m; row += warps) {
        const int8_t *p = a + static_cast<size_t>(row) * lda;
        int acc = 0;
        for (int c = lane * 16; c < k16; c += 32 * 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + c));
            acc = __dp4a(static_cast<int>(v.x), 0x01010101, acc);
            acc = __dp4a(static_cast<int>(v.y), 0x01010101, acc);
            acc = __dp4a(static_cast<int>(v.z), 0x01010101, acc);
            acc = __dp4a(static_cast<int>(v.w), 0x01010101, acc);
        }
        for (int c = k16 + lane; c < k; c += 32) acc += p[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) rs[row] = acc - zp_in * k;
    }
}
}  // namespace b200

extern "C" int b200_rowsum_i8(const void *a, int32_t lda, int32_t m, int32_t k, int32_t zp_in, int32_t *rs, void *stream)
{
    if (!a || !rs || m <= 0 || k <= 0 || lda < k || lda % 16 || (reinterpret_cast<uintptr_t>(a) & 15)) {
        b200::set_error("b200_rowsum_i8: bad arguments (m=%d k=%d lda=%d)", m, k, lda);
        return B200_ERR_ARG;
    }
    long long blocks = (static_cast<long long>(m) + 7) / 8;
    const long long cap = static_cast<long long>(b200::sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    b200::launch_kernel(b200::rowsum_i8_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, (cudaStream_t)stream,
                        static_cast<const int8_t *>(a), lda, m, k, zp_in, rs);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
