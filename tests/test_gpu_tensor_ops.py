"""GPU parity of transpose / gather / reduce_sum / layer_norm / rms_norm / matmul (csrc/tensor_ops.cu)
through the CSI-NN2 API: int8 bit-exact against the numpy oracle (tests/oracle_np.py, pinned against the reference in
tests/test_tensor_ops_oracle.py), fp16 within 1e-3 relative."""
import numpy as np
import pytest

import oracle_np as onp
from shl import (DT_F16, DT_INT8, H_GATHER, H_LAYER_NORM, H_MATMUL, H_REDUCE_SUM, H_RELU, H_RMS_NORM, H_TRANSPOSE, RM_GRAPH,
                 RM_LAYER, Layer)
from test_tensor_ops_oracle import TENSOR_CASES

pytestmark = pytest.mark.gpu
MODES = pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])


def _q(rng, shape):
    return rng.integers(-128, 128, size=shape, dtype=np.int8)


def f16_close(got, want, tol=1e-3):
    g, w = got.astype(np.float32), np.asarray(want, np.float32)
    err = np.abs(g - w) / np.maximum(np.abs(w), 1.0)
    assert err.max() <= tol, f"max relative error {err.max():.3e} > {tol}"


@MODES
@pytest.mark.parametrize("shape,perm", TENSOR_CASES["transpose"] + [((3, 64, 7, 9), (0, 2, 3, 1)), ((2, 40, 33), (0, 2, 1))])
def test_transpose(shape, perm, mode, b200, rng):
    x = _q(rng, shape)
    out_shape = tuple(shape[p] for p in perm)
    p4 = list(perm) + [0] * (4 - len(perm))
    layer = Layer(H_TRANSPOSE, out_shape, s_out=0.031, zp_out=4, kernel=p4[:2], stride=p4[2:])
    got = b200.run(DT_INT8, shape, [layer], x, s_in=0.05, zp_in=-3, run_mode=mode)
    assert np.array_equal(got, onp.transpose_i8(x, perm, 0.05, -3, 0.031, 4))
    same = Layer(H_TRANSPOSE, out_shape, s_out=0.05, zp_out=-3, kernel=p4[:2], stride=p4[2:])
    assert np.array_equal(b200.run(DT_INT8, shape, [same], x, s_in=0.05, zp_in=-3, run_mode=mode), np.transpose(x, perm))
    xh = rng.standard_normal(shape).astype(np.float16)
    assert np.array_equal(b200.run(DT_F16, shape, [Layer(H_TRANSPOSE, out_shape, kernel=p4[:2], stride=p4[2:])], xh,
                                   run_mode=mode).view(np.uint16), np.transpose(xh, perm).view(np.uint16))


@MODES
@pytest.mark.parametrize("shape,axis,idx", TENSOR_CASES["gather"])
def test_gather(shape, axis, idx, mode, b200, rng):
    x = _q(rng, shape)
    out_shape = shape[:axis] + (len(idx),) + shape[axis + 1:]
    layer = Layer(H_GATHER, out_shape, s_out=0.07, zp_out=9, w=np.asarray(idx, np.int64), axis=axis)
    got = b200.run(DT_INT8, shape, [layer], x, s_in=0.05, zp_in=-3, run_mode=mode)
    assert np.array_equal(got, onp.gather_i8(x, idx, axis, 0.05, -3, 0.07, 9))
    xh = rng.standard_normal(shape).astype(np.float16)
    goth = b200.run(DT_F16, shape, [Layer(H_GATHER, out_shape, w=np.asarray(idx, np.int64), axis=axis)], xh, run_mode=mode)
    assert np.array_equal(goth.view(np.uint16), onp.gather(xh, idx, axis).view(np.uint16))


@MODES
@pytest.mark.parametrize("shape,axis,keep", TENSOR_CASES["reduce_sum"] + [((2, 48, 9, 11), 1, True)])
def test_reduce_sum(shape, axis, keep, mode, b200, rng):
    x = _q(rng, shape)
    want_f = onp.reduce_sum_f32(onp.dequant(x, 0.02, 5), axis)
    out_shape = (1,) if axis == -1 else (shape[:axis] + ((1,) if keep else ()) + shape[axis + 1:])
    s_out = float(np.abs(want_f).max() / 100.0)
    layer = Layer(H_REDUCE_SUM, out_shape, s_out=s_out, zp_out=-2, axis=axis)
    got = b200.run(DT_INT8, shape, [layer], x, s_in=0.02, zp_in=5, run_mode=mode)
    assert np.array_equal(got, onp.quant(want_f, s_out, -2).reshape(out_shape))
    xh = rng.standard_normal(shape).astype(np.float16)
    goth = b200.run(DT_F16, shape, [Layer(H_REDUCE_SUM, out_shape, axis=axis)], xh, run_mode=mode)
    f16_close(goth, onp.reduce_sum_f32(xh.astype(np.float32), axis).reshape(out_shape), 2e-3 if axis == -1 else 1e-3)


@MODES
@pytest.mark.parametrize("shape,axis", TENSOR_CASES["norm"] + [((4, 7, 768), 2)])
@pytest.mark.parametrize("kind", [H_LAYER_NORM, H_RMS_NORM])
def test_layer_and_rms_norm(shape, axis, kind, mode, b200, rng):
    x = _q(rng, shape)
    n = int(np.prod(shape[axis:]))
    g = rng.integers(-100, 101, size=n, dtype=np.int8)
    b = rng.integers(-100, 101, size=n, dtype=np.int8)
    xf, gf, bf = onp.dequant(x, 0.04, 3), onp.dequant(g, 0.01, -2), onp.dequant(b, 0.01, -2)
    want_f = onp.layer_norm_f32(xf, gf, bf, axis, 1e-5) if kind == H_LAYER_NORM else onp.rms_norm_f32(xf, gf, axis, 1e-5)
    layer = Layer(kind, shape, s_out=0.03, zp_out=-5, w=g, b=b if kind == H_LAYER_NORM else None, s_w=np.float32([0.01]),
                  zp_w=np.int32([-2]), axis=axis, p0=1e-5)
    got = b200.run(DT_INT8, shape, [layer], x, s_in=0.04, zp_in=3, run_mode=mode)
    assert np.array_equal(got, onp.quant(want_f, 0.03, -5))
    xh = rng.standard_normal(shape).astype(np.float16)
    gh, bh = rng.standard_normal(n).astype(np.float16), rng.standard_normal(n).astype(np.float16)
    lh = Layer(kind, shape, w=gh, b=bh if kind == H_LAYER_NORM else None, axis=axis, p0=1e-5)
    wh = onp.layer_norm_f32(xh, gh, bh, axis, 1e-5) if kind == H_LAYER_NORM else onp.rms_norm_f32(xh, gh, axis, 1e-5)
    f16_close(b200.run(DT_F16, shape, [lh], xh, run_mode=mode), wh)


@MODES
@pytest.mark.parametrize("sa,sb,ta,tb", TENSOR_CASES["matmul"] + [((2, 3, 50, 96), (96, 200), False, False),
                                                                 ((4, 130, 64), (64, 64), False, False)])
def test_matmul_with_a_constant_operand(sa, sb, ta, tb, mode, b200, rng):
    """csinn_matmul on the tcgen05 GEMM: rows of mat0 packed K-major, mat1 packed once like fullyconnected weights with
    its zero point folded (asymmetric epilogue + row sums), output rows scattered back"""
    a, b = _q(rng, sa), _q(rng, sb)
    k = sa[-2] if ta else sa[-1]
    i = sa[-1] if ta else sa[-2]
    j = sb[-2] if tb else sb[-1]
    out_shape = sa[:-2] + (i, j)
    s_out = float(0.05 * 0.02 * np.sqrt(k) * 74 * 74 / 40.0)
    for zp_b in (6, 0):
        layer = Layer(H_MATMUL, out_shape, s_out=s_out, zp_out=2, w=b, s_w=np.float32([0.02]), zp_w=np.int32([zp_b]),
                      pad=(0, 0, int(ta), int(tb)))
        got = b200.run(DT_INT8, sa, [layer], a, s_in=0.05, zp_in=-4, run_mode=mode)
        want = onp.matmul_i8(a, b, ta, tb, 0.05, -4, 0.02, zp_b, s_out, 2)
        assert np.array_equal(got, want), (zp_b, np.count_nonzero(got != want), got.size)
    ah = rng.standard_normal(sa).astype(np.float16)
    bh = (rng.standard_normal(sb) / np.sqrt(k)).astype(np.float16)
    goth = b200.run(DT_F16, sa, [Layer(H_MATMUL, out_shape, w=bh, pad=(0, 0, int(ta), int(tb)))], ah, run_mode=mode)
    f16_close(goth, onp.matmul_f32(ah, bh, ta, tb))


@pytest.mark.parametrize("sa,tb", [((3, 20, 32), True), ((2, 2, 17, 40), True), ((4, 24, 24), False), ((1, 130, 64), True)])
def test_matmul_of_two_activations_fp16(sa, tb, b200, rng):
    """attention-style matmul: both operands are activations (fp16, batched): a x relu(a)^T (or a x relu(a) for square
    matrices) in graph mode -- the second operand is another tensor of the graph, packed K-major per run"""
    a = rng.standard_normal(sa).astype(np.float16)
    j = sa[-2] if tb else sa[-1]
    out_shape = sa[:-2] + (sa[-2], j)
    layers = [Layer(H_RELU, sa, in0=0), Layer(H_MATMUL, out_shape, in0=0, in1=1, pad=(0, 0, 0, int(tb)))]
    got = b200.run(DT_F16, sa, layers, a, run_mode=RM_GRAPH)
    f16_close(got, onp.matmul_f32(a, np.maximum(a, 0), False, tb))
