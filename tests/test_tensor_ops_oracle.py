"""CPU: the numpy restatements of transpose / gather / reduce_sum / layer_norm / rms_norm / matmul (tests/oracle_np.py)
pinned against the unmodified reference library through the public API."""
import numpy as np
import pytest

import oracle_np as onp
from shl import (DT_F16, DT_F32, DT_INT8, H_GATHER, H_LAYER_NORM, H_MATMUL, H_REDUCE_SUM, H_RMS_NORM, H_TRANSPOSE, Layer)

TENSOR_CASES = {
    "transpose": [((2, 5, 3, 4), (0, 2, 3, 1)), ((3, 4, 5), (2, 0, 1)), ((6, 7), (1, 0)), ((2, 3, 4, 5), (3, 1, 0, 2))],
    "gather": [((4, 6, 5), 1, [0, 5, -1, 2, 2, 9]), ((7, 3), 0, [6, 0, -7, 3]), ((2, 3, 4, 5), 3, [4, 1, -2])],
    "reduce_sum": [((2, 6, 4, 5), 1, False), ((3, 7, 5), 2, True), ((4, 9), 0, False), ((2, 3, 4), -1, False)],
    "norm": [((2, 5, 16), 2), ((3, 4, 6, 8), 2), ((6, 40), 1)],
    # the reference broadcasts a batch-less second operand only without transposes (source/reference/matmul.c:103-123)
    "matmul": [((3, 5, 8), (8, 6), False, False), ((7, 9), (5, 9), False, True), ((6, 10), (10, 4), False, False),
               ((9, 5), (9, 7), True, False), ((2, 4, 7, 9), (9, 5), False, False), ((12, 5), (4, 12), True, True)],
}


def _q(rng, shape):
    return rng.integers(-128, 128, size=shape, dtype=np.int8)


@pytest.mark.parametrize("shape,perm", TENSOR_CASES["transpose"])
def test_transpose_against_reference(shape, perm, ref, rng):
    x = _q(rng, shape)
    out_shape = tuple(shape[p] for p in perm)
    p4 = list(perm) + [0] * (4 - len(perm))
    layer = Layer(H_TRANSPOSE, out_shape, s_out=0.031, zp_out=4, kernel=p4[:2], stride=p4[2:])
    got = ref.run(DT_INT8, shape, [layer], x, s_in=0.05, zp_in=-3)
    assert np.array_equal(got, onp.transpose_i8(x, perm, 0.05, -3, 0.031, 4))


@pytest.mark.parametrize("shape,axis,idx", TENSOR_CASES["gather"])
def test_gather_against_reference(shape, axis, idx, ref, rng):
    x = _q(rng, shape)
    out_shape = shape[:axis] + (len(idx),) + shape[axis + 1:]
    layer = Layer(H_GATHER, out_shape, s_out=0.05, zp_out=-3, w=np.asarray(idx, np.int64), axis=axis)
    got = ref.run(DT_INT8, shape, [layer], x, s_in=0.05, zp_in=-3)
    assert np.array_equal(got, onp.gather_i8(x, idx, axis, 0.05, -3, 0.05, -3))


@pytest.mark.parametrize("shape,axis,keep", TENSOR_CASES["reduce_sum"])
def test_reduce_sum_against_reference(shape, axis, keep, ref, rng):
    x = _q(rng, shape)
    want_f = onp.reduce_sum_f32(onp.dequant(x, 0.02, 5), axis)
    if axis == -1:
        out_shape = (1,)
    elif keep:
        out_shape = shape[:axis] + (1,) + shape[axis + 1:]
    else:
        out_shape = shape[:axis] + shape[axis + 1:]
    s_out = float(np.abs(want_f).max() / 100.0)
    layer = Layer(H_REDUCE_SUM, out_shape, s_out=s_out, zp_out=-2, axis=axis)
    got = ref.run(DT_INT8, shape, [layer], x, s_in=0.02, zp_in=5)
    assert np.array_equal(got, onp.quant(want_f, s_out, -2).reshape(out_shape))


@pytest.mark.parametrize("shape,axis", TENSOR_CASES["norm"])
@pytest.mark.parametrize("kind", [H_LAYER_NORM, H_RMS_NORM])
def test_norm_against_reference(shape, axis, kind, ref_noavx, ref, rng):
    """the scalar build is matched bit for bit; the AVX build may contract `t / std * gamma + beta` into an fma
    (gcc -mfma): one LSB on a rare tie is tolerated there"""
    x = _q(rng, shape)
    n = int(np.prod(shape[axis:]))
    g = rng.integers(-100, 101, size=n, dtype=np.int8)
    b = rng.integers(-100, 101, size=n, dtype=np.int8)
    xf = onp.dequant(x, 0.04, 3)
    gf, bf = onp.dequant(g, 0.01, -2), onp.dequant(b, 0.01, -2)
    want_f = onp.layer_norm_f32(xf, gf, bf, axis, 1e-5) if kind == H_LAYER_NORM else onp.rms_norm_f32(xf, gf, axis, 1e-5)
    layer = Layer(kind, shape, s_out=0.03, zp_out=-5, w=g, b=b if kind == H_LAYER_NORM else None, s_w=np.float32([0.01]),
                  zp_w=np.int32([-2]), axis=axis, p0=1e-5)
    want = onp.quant(want_f, 0.03, -5)
    got = ref_noavx.run(DT_INT8, shape, [layer], x, s_in=0.04, zp_in=3)
    assert np.array_equal(got, want)
    d = np.abs(ref.run(DT_INT8, shape, [layer], x, s_in=0.04, zp_in=3).astype(int) - want.astype(int))
    assert d.max() <= 1 and np.count_nonzero(d) <= max(2, d.size // 500)


@pytest.mark.parametrize("sa,sb,ta,tb", TENSOR_CASES["matmul"])
def test_matmul_against_reference(sa, sb, ta, tb, ref, rng):
    a, b = _q(rng, sa), _q(rng, sb)
    k = sa[-2] if ta else sa[-1]
    i = sa[-1] if ta else sa[-2]
    j = sb[-2] if tb else sb[-1]
    out_shape = sa[:-2] + (i, j)
    s_out = float(0.05 * 0.02 * np.sqrt(k) * 74 * 74 / 40.0)
    layer = Layer(H_MATMUL, out_shape, s_out=s_out, zp_out=2, w=b, s_w=np.float32([0.02]), zp_w=np.int32([6]),
                  pad=(0, 0, int(ta), int(tb)))
    want = onp.matmul_i8(a, b, ta, tb, 0.05, -4, 0.02, 6, s_out, 2)
    got = ref.run(DT_INT8, sa, [layer], a, s_in=0.05, zp_in=-4)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and np.count_nonzero(d) <= max(2, d.size // 500), (d.max(), np.count_nonzero(d), d.size)
