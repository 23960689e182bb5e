// tensor_ops.cu -- the structural / normalisation operators of the RVV registration table that are not on the
// convolution hot path (source/thead_rvv/setup.c:154-508: transpose, gather, reduce_sum, layer_norm, rms_norm) and the
// row packing that lets csinn_matmul run on the tcgen05 GEMM.
//
// The API's tensors are logical row-major arrays of rank 1..4; on the device they live pixel-major (b200nn.h):
// rank 4 (d0, d1, d2, d3) = (n, c, h, w), rank 3 = (n, c, w), rank 2 = (n, c), rank 1 = (c), element (n, c, h, w)
// at ((n * H + h) * W + w) * cp + c.  Every kernel here walks LOGICAL indices -- the order the reference's loops
// define their semantics in -- and maps them to device offsets, one thread per output element (or per
// normalisation row).  They are bandwidth- or latency-bound helpers, written for exactness, not for speed:
//   * int8 copies requantise through a 256-entry table requant_out(dequant_in(q)) like concat / split;
//   * sums are sequential f32 in the reference's order (source/reference/reduce_sum.c:21, layer_norm.c:21,
//     rms_norm.c:21), so the int8 results match the reference bit for bit.
#include "common.cuh"

namespace b200 {

struct View {
    int rank, d[4], cp;
};

__device__ __forceinline__ long long view_offset(const View &v, const int (&i)[4])
{
    int n = 0, c = 0, h = 0, w = 0, H = 1, W = 1;
    switch (v.rank) {
        case 4: n = i[0], c = i[1], h = i[2], w = i[3], H = v.d[2], W = v.d[3]; break;
        case 3: n = i[0], c = i[1], w = i[2], W = v.d[2]; break;
        case 2: n = i[0], c = i[1]; break;
        default: c = i[0]; break;
    }
    return ((static_cast<long long>(n) * H + h) * W + w) * v.cp + c;
}
__device__ __forceinline__ void view_unravel(const View &v, long long lin, int (&i)[4])
{
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        if (k < v.rank) {
            i[k] = static_cast<int>(lin % v.d[k]);
            lin /= v.d[k];
        } else {
            i[k] = 0;
        }
    }
}
__device__ __forceinline__ long long view_lin2off(const View &v, long long lin)
{
    int i[4];
    view_unravel(v, lin, i);
    return view_offset(v, i);
}
__device__ __forceinline__ long long view_size(const View &v)
{
    long long s = 1;
    for (int k = 0; k < v.rank; k++) s *= v.d[k];
    return s;
}

__device__ __forceinline__ float load_f(const void *p, long long off, int eb, float s, int zp)
{
    if (eb == 1) return dequant_i8(static_cast<const int8_t *>(p)[off], s, zp);
    return __half2float(static_cast<const __half *>(p)[off]);
}
__device__ __forceinline__ void store_f(void *p, long long off, int eb, float v, float s, int zp)
{
    if (eb == 1)
        static_cast<int8_t *>(p)[off] = static_cast<int8_t>(quant_i8_exact(v, s, zp));
    else
        static_cast<__half *>(p)[off] = __float2half_rn(v);
}

// ---- transpose: out[o0..] = in[i], i[perm[k]] = o[k] (source/reference/transpose.c:57-70) ----------------------
__global__ void __launch_bounds__(256) permute_kernel(View in, View out, int p0, int p1, int p2, int p3, const void *src,
                                                      void *dst, int eb, const int8_t *__restrict__ lut)
{
    pdl_launch_dependents();
    pdl_wait();
    const int perm[4] = {p0, p1, p2, p3};
    const long long total = view_size(out);
    for (long long lin = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; lin < total;
         lin += static_cast<long long>(gridDim.x) * blockDim.x) {
        int o[4], i[4] = {0, 0, 0, 0};
        view_unravel(out, lin, o);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < out.rank) i[perm[k]] = o[k];
        const long long so = view_offset(in, i), doff = view_offset(out, o);
        if (eb == 1) {
            const int8_t q = static_cast<const int8_t *>(src)[so];
            static_cast<int8_t *>(dst)[doff] = lut ? lut[static_cast<int>(q) + 128] : q;
        } else {
            static_cast<uint16_t *>(dst)[doff] = static_cast<const uint16_t *>(src)[so];
        }
    }
}

// ---- gather along one axis with constant indices (source/reference/gather.c:21-60) ---------------------------------
__global__ void __launch_bounds__(256) gather_kernel(View in, View out, int axis_dim, long long inner, int n_idx,
                                                     const int *__restrict__ idx, const void *src, void *dst, int eb,
                                                     const int8_t *__restrict__ lut, int oob_q)
{
    pdl_launch_dependents();
    pdl_wait();
    const long long total = view_size(out);
    for (long long lin = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; lin < total;
         lin += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long in_i = lin % inner, t = lin / inner;
        const int j = static_cast<int>(t % n_idx);
        const long long outer = t / n_idx;
        int id = idx[j];
        if (id < 0) id += axis_dim;
        const long long doff = view_lin2off(out, lin);
        if (id < 0 || id >= axis_dim) {  // the reference zero-fills (in the real domain)
            if (eb == 1)
                static_cast<int8_t *>(dst)[doff] = static_cast<int8_t>(oob_q);
            else
                static_cast<uint16_t *>(dst)[doff] = 0;
            continue;
        }
        const long long so = view_lin2off(in, (outer * axis_dim + id) * inner + in_i);
        if (eb == 1) {
            const int8_t q = static_cast<const int8_t *>(src)[so];
            static_cast<int8_t *>(dst)[doff] = lut ? lut[static_cast<int>(q) + 128] : q;
        } else {
            static_cast<uint16_t *>(dst)[doff] = static_cast<const uint16_t *>(src)[so];
        }
    }
}

// ---- reduce_sum over one axis, or over everything (axis < 0) (source/reference/reduce_sum.c:21-60) ----------------
__global__ void __launch_bounds__(128) reduce_sum_kernel(View in, View out, int axis, const void *src, void *dst, int eb,
                                                         float s_in, int zp_in, float s_out, int zp_out)
{
    pdl_launch_dependents();
    pdl_wait();
    long long inner = 1;
    for (int k = axis + 1; k < in.rank; k++) inner *= in.d[k];
    const int cnt = axis < 0 ? 1 : in.d[axis];
    const long long total = axis < 0 ? 1 : view_size(in) / cnt;
    for (long long lin = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; lin < total;
         lin += static_cast<long long>(gridDim.x) * blockDim.x) {
        float acc = 0.f;
        if (axis < 0) {
            const long long n = view_size(in);
            for (long long j = 0; j < n; j++) acc = __fadd_rn(acc, load_f(src, view_lin2off(in, j), eb, s_in, zp_in));
        } else {
            const long long outer = lin / inner, in_i = lin % inner;
            for (int j = 0; j < cnt; j++)
                acc = __fadd_rn(acc, load_f(src, view_lin2off(in, (outer * cnt + j) * inner + in_i), eb, s_in, zp_in));
        }
        store_f(dst, view_lin2off(out, lin), eb, acc, s_out, zp_out);
    }
}

// ---- layer_norm / rms_norm over the trailing axes (source/reference/layer_norm.c:21-66, rms_norm.c:21-52) ---------
// One thread per normalisation row: the reference's three sequential f32 passes, operation for operation (sqrt in
// double on the float sum, as the C expression sqrt(var + eps) / 1.0 / sqrt(...) evaluates).
__global__ void __launch_bounds__(64) norm_kernel(int rms, View v, long long batches, int norm_size, float eps,
                                                  const float *__restrict__ gamma, const float *__restrict__ beta,
                                                  const void *src, void *dst, int eb, float s_in, int zp_in, float s_out,
                                                  int zp_out)
{
    pdl_launch_dependents();
    pdl_wait();
    for (long long b = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; b < batches;
         b += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long base = b * norm_size;
        if (rms) {
            float sum = 0.f;
            for (int i = 0; i < norm_size; i++) {
                const float x = load_f(src, view_lin2off(v, base + i), eb, s_in, zp_in);
                sum = __fadd_rn(sum, __fmul_rn(x, x));
            }
            const float scale = static_cast<float>(1.0 / sqrt(static_cast<double>(__fadd_rn(__fdiv_rn(sum, static_cast<float>(norm_size)), eps))));
            for (int i = 0; i < norm_size; i++) {
                const long long off = view_lin2off(v, base + i);
                const float x = load_f(src, off, eb, s_in, zp_in);
                store_f(dst, off, eb, __fmul_rn(__fmul_rn(x, scale), gamma[i]), s_out, zp_out);
            }
        } else {
            float mean = 0.f;
            for (int i = 0; i < norm_size; i++) mean = __fadd_rn(mean, load_f(src, view_lin2off(v, base + i), eb, s_in, zp_in));
            mean = __fdiv_rn(mean, static_cast<float>(norm_size));
            float sum = 0.f;
            for (int i = 0; i < norm_size; i++) {
                const float t = __fsub_rn(load_f(src, view_lin2off(v, base + i), eb, s_in, zp_in), mean);
                sum = __fadd_rn(sum, __fmul_rn(t, t));
            }
            const float var = __fdiv_rn(sum, static_cast<float>(norm_size));
            const float sd = static_cast<float>(sqrt(static_cast<double>(__fadd_rn(var, eps))));
            for (int i = 0; i < norm_size; i++) {
                const long long off = view_lin2off(v, base + i);
                const float t = __fsub_rn(load_f(src, off, eb, s_in, zp_in), mean);
                store_f(dst, off, eb, __fadd_rn(__fmul_rn(__fdiv_rn(t, sd), gamma[i]), beta[i]), s_out, zp_out);
            }
        }
    }
}

// ---- matmul on the tensor cores: logical matrices <-> dense K-major rows --------------------------------------------
// pack: rows[(b * R + r)][k] = t[b][r][k] (trans = 0, the tensor is [.., R, K]) or t[b][k][r] (trans = 1, [.., K, R]),
// zero-padded to the row pitch ld; unpack: t[b][r][c] = rows[(b * R + r)][c].
__global__ void __launch_bounds__(256) pack_rows_kernel(View t, int batches, int R, int K, int trans, const void *src,
                                                        void *rows, int ld, int eb)
{
    pdl_launch_dependents();
    pdl_wait();
    const long long total = static_cast<long long>(batches) * R * ld;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % ld);
        const long long br = i / ld;
        const int r = static_cast<int>(br % R);
        const long long b = br / R;
        if (k >= K) {
            if (eb == 1) static_cast<int8_t *>(rows)[i] = 0;
            else static_cast<uint16_t *>(rows)[i] = 0;
            continue;
        }
        const long long lin = trans ? (b * K + k) * R + r : (b * R + r) * K + k;
        const long long so = view_lin2off(t, lin);
        if (eb == 1) static_cast<int8_t *>(rows)[i] = static_cast<const int8_t *>(src)[so];
        else static_cast<uint16_t *>(rows)[i] = static_cast<const uint16_t *>(src)[so];
    }
}
__global__ void __launch_bounds__(256) unpack_rows_kernel(View t, long long nrows, int C, const void *rows, int ld, void *dst,
                                                          int eb)
{
    pdl_launch_dependents();
    pdl_wait();
    const long long total = nrows * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / C;
        const int c = static_cast<int>(i % C);
        const long long doff = view_lin2off(t, i);
        if (eb == 1) static_cast<int8_t *>(dst)[doff] = static_cast<const int8_t *>(rows)[r * ld + c];
        else static_cast<uint16_t *>(dst)[doff] = static_cast<const uint16_t *>(rows)[r * ld + c];
    }
}

static bool to_view(const b200_view *b, View *v)
{
    if (!b || b->rank < 1 || b->rank > 4 || b->cp <= 0) return false;
    v->rank = b->rank, v->cp = b->cp;
    for (int k = 0; k < 4; k++) v->d[k] = k < b->rank ? b->dim[k] : 1;
    for (int k = 0; k < b->rank; k++)
        if (b->dim[k] <= 0) return false;
    return true;
}
static long long host_size(const View &v)
{
    long long s = 1;
    for (int k = 0; k < v.rank; k++) s *= v.d[k];
    return s;
}
static int grid_for(long long n, int block, int per_sm)
{
    long long g = (n + block - 1) / block;
    const long long cap = static_cast<long long>(sm_count()) * per_sm;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b200

using namespace b200;

extern "C" int b200_permute(const b200_view *in, const void *src, const b200_view *out, void *dst, const int32_t *perm,
                            int elem_bytes, const int8_t *lut_dev, void *stream)
{
    View vi, vo;
    if (!to_view(in, &vi) || !to_view(out, &vo) || !src || !dst || !perm || vi.rank != vo.rank || (elem_bytes != 1 && elem_bytes != 2)) {
        set_error("b200_permute: bad arguments");
        return B200_ERR_ARG;
    }
    int p[4] = {0, 1, 2, 3};
    for (int k = 0; k < vo.rank; k++) {
        p[k] = perm[k];
        if (p[k] < 0 || p[k] >= vi.rank || vi.d[p[k]] != vo.d[k]) {
            set_error("b200_permute: permutation does not map the input shape onto the output shape");
            return B200_ERR_ARG;
        }
    }
    launch_kernel(permute_kernel, dim3(grid_for(host_size(vo), 256, 16)), dim3(256), 0, (cudaStream_t)stream, vi, vo, p[0], p[1],
                  p[2], p[3], src, dst, elem_bytes, lut_dev);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_gather(const b200_view *in, const void *src, const b200_view *out, void *dst, int axis,
                           const int32_t *idx_dev, int n_idx, int elem_bytes, const int8_t *lut_dev, int oob_q, void *stream)
{
    View vi, vo;
    if (!to_view(in, &vi) || !to_view(out, &vo) || !src || !dst || !idx_dev || n_idx <= 0 || axis < 0 || axis >= vi.rank ||
        (elem_bytes != 1 && elem_bytes != 2)) {
        set_error("b200_gather: bad arguments");
        return B200_ERR_ARG;
    }
    long long inner = 1, outer = 1;
    for (int k = axis + 1; k < vi.rank; k++) inner *= vi.d[k];
    for (int k = 0; k < axis; k++) outer *= vi.d[k];
    if (host_size(vo) != outer * n_idx * inner) {
        set_error("b200_gather: output has %lld elements, outer x indices x inner = %lld", host_size(vo), outer * n_idx * inner);
        return B200_ERR_ARG;
    }
    launch_kernel(gather_kernel, dim3(grid_for(host_size(vo), 256, 16)), dim3(256), 0, (cudaStream_t)stream, vi, vo, vi.d[axis],
                  inner, n_idx, reinterpret_cast<const int *>(idx_dev), src, dst, elem_bytes, lut_dev, oob_q);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_reduce_sum(const b200_view *in, const void *src, const b200_view *out, void *dst, int axis, int elem_bytes,
                               float s_in, int zp_in, float s_out, int zp_out, void *stream)
{
    View vi, vo;
    if (!to_view(in, &vi) || !to_view(out, &vo) || !src || !dst || axis >= vi.rank || (elem_bytes != 1 && elem_bytes != 2)) {
        set_error("b200_reduce_sum: bad arguments");
        return B200_ERR_ARG;
    }
    const long long want = axis < 0 ? 1 : host_size(vi) / vi.d[axis];
    if (host_size(vo) != want) {
        set_error("b200_reduce_sum: output has %lld elements, expected %lld", host_size(vo), want);
        return B200_ERR_ARG;
    }
    launch_kernel(reduce_sum_kernel, dim3(grid_for(want, 128, 16)), dim3(128), 0, (cudaStream_t)stream, vi, vo, axis, src, dst,
                  elem_bytes, s_in, zp_in, s_out, zp_out);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_norm(int rms, const b200_view *v, const void *src, void *dst, int axis, float eps, const float *gamma_dev,
                         const float *beta_dev, int elem_bytes, float s_in, int zp_in, float s_out, int zp_out, void *stream)
{
    View vv;
    if (!to_view(v, &vv) || !src || !dst || !gamma_dev || (!rms && !beta_dev) || axis < 0 || axis >= vv.rank ||
        (elem_bytes != 1 && elem_bytes != 2)) {
        set_error("b200_norm: bad arguments");
        return B200_ERR_ARG;
    }
    long long batches = 1, norm = 1;
    for (int k = 0; k < axis; k++) batches *= vv.d[k];
    for (int k = axis; k < vv.rank; k++) norm *= vv.d[k];
    if (norm >= (1ll << 31)) {
        set_error("b200_norm: normalised extent too large");
        return B200_ERR_UNSUPPORTED;
    }
    launch_kernel(norm_kernel, dim3(grid_for(batches, 64, 32)), dim3(64), 0, (cudaStream_t)stream, rms, vv, batches,
                  static_cast<int>(norm), eps, gamma_dev, beta_dev, src, dst, elem_bytes, s_in, zp_in, s_out, zp_out);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_pack_rows(const b200_view *t, const void *src, int batches, int rows, int k, int trans, void *rows_dev,
                              int ld, int elem_bytes, void *stream)
{
    View v;
    if (!to_view(t, &v) || !src || !rows_dev || batches <= 0 || rows <= 0 || k <= 0 || ld < k ||
        host_size(v) != static_cast<long long>(batches) * rows * k) {
        set_error("b200_pack_rows: bad arguments");
        return B200_ERR_ARG;
    }
    launch_kernel(pack_rows_kernel, dim3(grid_for(static_cast<long long>(batches) * rows * ld, 256, 16)), dim3(256), 0,
                  (cudaStream_t)stream, v, batches, rows, k, trans, src, rows_dev, ld, elem_bytes);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_unpack_rows(const b200_view *t, void *dst, long long nrows, int cols, const void *rows_dev, int ld,
                                int elem_bytes, void *stream)
{
    View v;
    if (!to_view(t, &v) || !dst || !rows_dev || nrows <= 0 || cols <= 0 || ld < cols || host_size(v) != nrows * cols) {
        set_error("b200_unpack_rows: bad arguments");
        return B200_ERR_ARG;
    }
    launch_kernel(unpack_rows_kernel, dim3(grid_for(nrows * cols, 256, 16)), dim3(256), 0, (cudaStream_t)stream, v, nrows, cols,
                  rows_dev, ld, dst, elem_bytes);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
