"""TEST INFRASTRUCTURE ONLY -- numpy restatements of the structural / normalisation operators of the reference
(transpose, gather, reduce_sum, layer_norm, rms_norm, matmul), each following the reference's float sequence line by
line; pinned against the unmodified reference library in tests/test_oracle.py, then used as the checker of the CUDA
kernels in tests/test_gpu_tensor_ops.py.  Never imported by the product.

int8 tensors are dequantised / requantised exactly as the reference does around its f32 kernels
(shl_ref_siso_callback_base, source/reference/utils.c:609): (q - zp) * s in f32, and
clamp(nearbyint(x / s) + zp) with an f32 division (source/nn2/utils.c:550 float_to_int8_base).
"""
import numpy as np

F = np.float32


def dequant(q, s, zp):
    return (q.astype(F) - F(zp)) * F(s)


def quant(x, s, zp):
    r = np.rint((x.astype(F) / F(s)).astype(np.float64)) + float(zp)
    return np.clip(r.astype(F), -128, 127).astype(np.int8)


def transpose_i8(x, perm, s_in, zp_in, s_out, zp_out):
    """source/reference/transpose.c:57-70 between shl_ref_tensor_transform_f32 and csinn_tensor_data_convert"""
    return quant(np.transpose(dequant(x, s_in, zp_in), perm), s_out, zp_out)


def gather(x, idx, axis):
    """source/reference/gather.c:21-60: negative indices count from the end, out-of-range ones give 0.0"""
    n = x.shape[axis]
    idx = np.asarray(idx, np.int64)
    fixed = np.where(idx < 0, idx + n, idx)
    ok = (fixed >= 0) & (fixed < n)
    out = np.take(x, np.where(ok, fixed, 0), axis=axis)
    shape = [1] * out.ndim
    shape[axis] = idx.size
    return np.where(ok.reshape(shape), out, np.zeros((), x.dtype))


def gather_i8(x, idx, axis, s_in, zp_in, s_out, zp_out):
    return quant(gather(dequant(x, s_in, zp_in), idx, axis), s_out, zp_out)


def reduce_sum_f32(x, axis):
    """source/reference/reduce_sum.c:21-60: sequential f32 sum along the axis (axis -1: over everything, in memory order)"""
    x = x.astype(F)
    if axis == -1:
        acc = F(0)
        for v in x.reshape(-1):
            acc = F(acc + v)
        return np.asarray([acc], F)
    acc = np.zeros(x.shape[:axis] + x.shape[axis + 1:], F)
    for j in range(x.shape[axis]):
        acc = (acc + np.take(x, j, axis=axis)).astype(F)
    return acc


def layer_norm_f32(x, gamma, beta, axis, eps):
    """source/reference/layer_norm.c:21-66: mean, variance and the affine map as three sequential f32 passes; sqrt of
    the f32 sum in double, narrowed to f32"""
    x = x.astype(F)
    lead = int(np.prod(x.shape[:axis], dtype=np.int64))
    rows = x.reshape(lead, -1)
    n = rows.shape[1]
    g, b = gamma.astype(F).reshape(-1), beta.astype(F).reshape(-1)
    mean = np.zeros(lead, F)
    for i in range(n):
        mean = (mean + rows[:, i]).astype(F)
    mean = (mean / F(n)).astype(F)
    t = (rows - mean[:, None]).astype(F)
    ssum = np.zeros(lead, F)
    for i in range(n):
        ssum = (ssum + (t[:, i] * t[:, i]).astype(F)).astype(F)
    var = (ssum / F(n)).astype(F)
    sd = np.sqrt((var + F(eps)).astype(F).astype(np.float64)).astype(F)
    out = (((t / sd[:, None]).astype(F) * g[None, :]).astype(F) + b[None, :]).astype(F)
    return out.reshape(x.shape)


def rms_norm_f32(x, weight, axis, eps):
    """source/reference/rms_norm.c:21-52: scale = 1.0 / sqrt(sum / n + eps) evaluated in double on f32 operands"""
    x = x.astype(F)
    lead = int(np.prod(x.shape[:axis], dtype=np.int64))
    rows = x.reshape(lead, -1)
    n = rows.shape[1]
    ssum = np.zeros(lead, F)
    for i in range(n):
        ssum = (ssum + (rows[:, i] * rows[:, i]).astype(F)).astype(F)
    scale = (1.0 / np.sqrt(((ssum / F(n)).astype(F) + F(eps)).astype(F).astype(np.float64))).astype(F)
    out = ((rows * scale[:, None]).astype(F) * weight.astype(F).reshape(1, -1)).astype(F)
    return out.reshape(x.shape)


def matmul_i8(a, b, trans_a, trans_b, s_a, zp_a, s_b, zp_b, s_out, zp_out):
    """csinn_matmul on int8 in the integer form of the requantise contract (include/b200nn.h): exact int32
    sum of (a - zp_a) * (b - zp_b), one fmaf with mult = (float)((double)s_a * s_b / s_out), round half even.  The
    reference accumulates dequantised f32 products (source/reference/matmul.c:21-60) and so differs by +-1 LSB on a few
    ppm of outputs, like the convolutions."""
    A = a.astype(np.int64) - zp_a
    B = b.astype(np.int64) - zp_b
    if trans_a:
        A = np.swapaxes(A, -1, -2)
    if trans_b:
        B = np.swapaxes(B, -1, -2)
    acc = np.matmul(A, B)
    mult = F(np.float64(F(s_a)) * np.float64(F(s_b)) / np.float64(F(s_out)))
    f = (acc.astype(np.float64) * np.float64(mult))  # |acc| < 2^24: float(acc) is exact and the product fits a double
    r = np.rint(f.astype(F).astype(np.float64)) + zp_out
    return np.clip(r, -128, 127).astype(np.int8)


def matmul_f32(a, b, trans_a, trans_b):
    A, B = a.astype(F), b.astype(F)
    if trans_a:
        A = np.swapaxes(A, -1, -2)
    if trans_b:
        B = np.swapaxes(B, -1, -2)
    return np.matmul(A.astype(np.float64), B.astype(np.float64)).astype(F)
