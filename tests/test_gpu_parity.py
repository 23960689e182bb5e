"""GPU parity tests proper (-m gpu): every case goes through the CSI-NN2 public API into
libshl_b200.so (registry -> b200_opt -> C-ABI shim -> sm_100a kernels) and is compared with the
oracle on the same seeded inputs.

Bar: int8 results are BIT-EXACT against oracle/oracle_int.c (the arithmetic contract in
include/b200nn.h) and within the reference's own f32-noise band (|d| <= 1 on <= 1e-4 of outputs,
see tests/test_oracle.py) against the unmodified reference library; fp16 is within 1e-3 relative
(|d| / max(|want|, 1), the north-star tolerance) of f32-accumulated math.
"""
import os

import numpy as np
import pytest

import nets
from shl import (ACT_NONE, ACT_RELU, ACT_RELU6, API_C906, API_C920, API_RVV, DT_F16, DT_INT8, H_ADD, H_AVGPOOL,
                 H_CONV, H_CONV_RELU, H_CONV_RELU6, H_DWCONV, H_FC, H_FLATTEN, H_GAP, H_MAXPOOL, H_RELU, H_RELU6,
                 H_SOFTMAX, RM_GRAPH, RM_LAYER, Layer, conv_out_hw, synth_conv_i8)

pytestmark = pytest.mark.gpu
F16_TOL = 1e-3


def f16_close(got, want, tol=F16_TOL):
    g, w = got.astype(np.float32), np.asarray(want, np.float32)
    err = np.abs(g - w) / np.maximum(np.abs(w), 1.0)
    assert err.max() <= tol, f"max relative error {err.max():.3e} > {tol}"


def ref_band(got, want):
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    # measured (tests/test_oracle.py prints it): the reference's f32 accumulation flips 2-6e-5 of the outputs by one
    # LSB against exact integer accumulation; the gate is 1e-4 (at least 2 outputs on small tensors)
    rate = np.count_nonzero(d) / d.size
    assert d.max() <= 1 and np.count_nonzero(d) <= max(1e-4 * d.size, 2), (d.max(), np.count_nonzero(d), d.size, rate)


CONV_CASES = [
    # n, c, h, w, o, k, stride, pad, group, depthwise, zp_in, kind
    (1, 128, 8, 16, 64, 1, 1, 0, 1, False, 0, H_CONV),       # one GEMM tile
    (1, 32, 8, 16, 64, 1, 1, 0, 1, False, 0, H_CONV),        # K shorter than a k-block
    (1, 16, 8, 16, 16, 1, 1, 0, 1, False, 0, H_CONV),        # smallest tile
    (1, 64, 10, 10, 64, 1, 1, 0, 1, False, 0, H_CONV),       # ragged M
    (1, 64, 8, 16, 40, 1, 1, 0, 1, False, 0, H_CONV),        # ragged N
    (1, 72, 8, 16, 64, 1, 1, 0, 1, False, 0, H_CONV),        # ragged K
    (1, 20, 9, 9, 24, 1, 1, 0, 1, False, -3, H_CONV),        # channels not a multiple of 16
    (1, 1024, 7, 7, 128, 1, 1, 0, 1, False, -128, H_CONV),   # 8 k-blocks, extreme zero point
    (2, 1024, 1, 1, 1000, 1, 1, 0, 1, False, 0, H_CONV),     # MobileNetV1 classifier: 4 n-tiles
    (2, 128, 56, 56, 128, 1, 1, 0, 1, False, 0, H_CONV),     # many tiles per CTA (double-buffered TMEM)
    (2, 512, 14, 14, 512, 1, 1, 0, 1, False, -128, H_CONV),  # 256-column n-tiles, resident weights, 4 k-blocks
    (1, 128, 28, 28, 256, 1, 1, 0, 1, False, 0, H_CONV),     # one 256-column n-tile: activations read once
    (3, 96, 9, 11, 200, 1, 1, 0, 1, False, 7, H_CONV),       # ragged 256-column tile (N = 200), ragged M and K
    (1, 256, 20, 20, 400, 1, 1, 0, 1, False, 0, H_CONV_RELU),  # two n-tiles, the second ragged
    (1, 64, 14, 14, 64, 1, 1, 0, 1, False, 0, H_CONV_RELU),
    (1, 64, 14, 14, 64, 1, 1, 0, 1, False, 3, H_CONV_RELU6),
    (1, 32, 14, 14, 48, 3, 1, 1, 1, False, -7, H_CONV),      # im2col path, asymmetric pad value
    (2, 3, 32, 32, 32, 3, 2, 1, 1, False, 5, H_CONV),        # first-layer shape
    (1, 3, 40, 40, 64, 7, 2, 3, 1, False, 0, H_CONV),        # ResNet stem shape
    (1, 64, 13, 13, 64, 1, 2, 0, 1, False, 0, H_CONV),       # strided 1x1 (ResNet downsample)
    (1, 64, 12, 12, 64, 3, 1, 1, 4, False, 0, H_CONV),       # group conv
    (2, 32, 11, 13, 96, 3, 1, 1, 2, False, -3, H_CONV),      # group conv, 48 outputs per group: a 64-column tile clipped
                                                             # at the group's window (it used to spill into the next group)
    (1, 60, 9, 9, 240, 1, 1, 0, 3, False, 5, H_CONV),        # grouped 1x1, 80 outputs per group (128-column tile)
    (1, 48, 10, 10, 320, 3, 2, 1, 2, False, 0, H_CONV_RELU), # 160 per group: one full and one clipped n-tile per group
    (2, 32, 20, 20, 32, 3, 1, 1, 1, True, -7, H_CONV),       # depthwise via csinn_conv2d
    (1, 64, 21, 21, 64, 3, 2, 1, 1, True, 0, H_DWCONV),      # depthwise via csinn_depthwise_conv2d
    (1, 24, 9, 11, 24, 3, 1, 1, 1, True, 2, H_CONV),         # depthwise, ragged channels / width
    (1, 16, 12, 12, 16, 5, 1, 2, 1, True, 4, H_CONV),        # depthwise 5x5
    (1, 512, 14, 14, 512, 3, 1, 1, 1, True, -128, H_CONV),   # MobileNetV1 depthwise shape
    # outputs per group that are not a multiple of 16 bytes go through a staging tensor and a slice copy per group
    (1, 48, 9, 10, 72, 3, 1, 1, 3, False, -5, H_CONV),       # 24 per group
    (2, 32, 7, 7, 80, 1, 1, 0, 2, False, 0, H_CONV_RELU),    # 40 per group, 1x1
    (1, 16, 11, 9, 32, 3, 1, 1, 16, False, 6, H_CONV),       # depthwise with depth multiplier 2 (group = C, O = 2C)
    (1, 8, 10, 10, 24, 3, 2, 1, 8, False, -3, H_DWCONV),     # depth multiplier 3 through csinn_depthwise_conv2d
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d_g%d_dw%d_zp%d_op%d" % c)
def test_conv_int8_bit_exact(case, b200, oracle, rng):
    n, c, h, w, o, k, stride, pad, group, dw, zp_in, kind = case
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k, group=group, depthwise=dw)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layer = Layer(kind, (n, o, oh, ow), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w, stride=(stride, stride),
                  pad=(pad,) * 4, group=c if dw else group)
    act = {H_CONV_RELU: ACT_RELU, H_CONV_RELU6: ACT_RELU6}.get(kind, ACT_NONE)
    got = b200.run(DT_INT8, (n, c, h, w), [layer], x, s_in=0.02, zp_in=zp_in)
    want = oracle.conv2d_i8(x, wt, b, (n, o, oh, ow), depthwise=dw, stride=(stride, stride), pad=(pad,) * 4,
                            dilation=(1, 1), group=group, s_in=0.02, zp_in=zp_in, s_w=s_w, s_b=None, s_out=s_out,
                            zp_out=3, act=act)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} outputs differ from the oracle"


@pytest.mark.parametrize("case", [
    # n, c, h, w, o, k, stride, pad, dilation, group, depthwise, zp_in
    (2, 32, 19, 17, 48, 3, 1, 2, 2, 1, False, -7),     # dilation 2, "same" padding (im2col + GEMM)
    (1, 16, 20, 20, 32, 3, 2, 3, 3, 1, False, 4),      # dilation 3 with stride 2
    (1, 3, 33, 31, 16, 3, 1, 2, 2, 1, False, -5),      # dilated first layer read from the NCHW graph input
    (2, 24, 15, 15, 24, 3, 1, 2, 2, 1, True, 6),       # dilated depthwise (generic kernel)
    (1, 64, 14, 14, 64, 3, 1, 4, 4, 4, False, 0),      # dilated group conv
], ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d_d%d_g%d_dw%d_zp%d" % c)
@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
def test_dilated_conv_int8_bit_exact(case, mode, b200, oracle, ref_noavx, rng):
    """dilation > 1 (source/reference/convolution.c:28-89 handles it; its AVX path ignores it,
    conv_avx.h:109, so the reference is consulted through the non-AVX build)"""
    n, c, h, w, o, k, stride, pad, dil, group, dw, zp_in = case
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k, group=group, depthwise=dw)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4, (dil, dil))
    layer = Layer(H_CONV, (n, o, oh, ow), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w, stride=(stride, stride),
                  pad=(pad,) * 4, dilation=(dil, dil), group=c if dw else group)
    got = b200.run(DT_INT8, (n, c, h, w), [layer], x, s_in=0.02, zp_in=zp_in, run_mode=mode)
    want = oracle.conv2d_i8(x, wt, b, (n, o, oh, ow), depthwise=dw, stride=(stride, stride), pad=(pad,) * 4,
                            dilation=(dil, dil), group=group, s_in=0.02, zp_in=zp_in, s_w=s_w, s_b=None, s_out=s_out,
                            zp_out=3)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} outputs differ from the oracle"
    if mode == RM_LAYER:
        ref_band(got, ref_noavx.run(DT_INT8, (n, c, h, w), [layer], x, s_in=0.02, zp_in=zp_in))


@pytest.mark.parametrize("case", [(1, 48, 9, 9, 24, 3, 1, 1, 3), (2, 16, 10, 11, 48, 1, 1, 0, 2), (1, 24, 8, 8, 120, 3, 1, 1, 3),
                                  (1, 16, 8, 8, 16, 3, 1, 1, 4), (1, 8, 9, 9, 16, 3, 1, 1, 8)],  # 4 per group (staged); depth multiplier 2
                         ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d_g%d" % c)
def test_group_conv_fp16_clips_tiles_at_the_group_window(case, b200, oracle, rng):
    """fp16 group conv whose outputs per group (8, 24, 40) are narrower than the GEMM's n-tile: every group's tile
    must be clipped at its own column window of the shared output"""
    n, c, h, w, o, k, stride, pad, group = case
    x = rng.standard_normal((n, c, h, w)).astype(np.float16)
    cg = c // group
    wt = (rng.standard_normal((o, cg, k, k)) / np.sqrt(cg * k * k)).astype(np.float16)
    b = rng.standard_normal(o).astype(np.float16)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layer = Layer(H_CONV, (n, o, oh, ow), w=wt, b=b, stride=(stride, stride), pad=(pad,) * 4, group=group)
    got = b200.run(DT_F16, (n, c, h, w), [layer], x)
    want = oracle.conv2d_f32(x.astype(np.float32), wt.astype(np.float32), b.astype(np.float32), (n, o, oh, ow),
                             stride=(stride, stride), pad=(pad,) * 4, group=group)
    f16_close(got, want)


def test_conv_int8_against_the_reference_library(b200, ref, rng):
    """the same API calls against the unmodified reference: inside its own f32-noise band"""
    for (n, c, h, w, o, k, stride, pad, dw, zp_in) in [(1, 128, 28, 28, 128, 1, 1, 0, False, 0),
                                                      (1, 64, 14, 14, 96, 3, 1, 1, False, -7),
                                                      (1, 64, 28, 28, 64, 3, 2, 1, True, -7)]:
        x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
        wt, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k, depthwise=dw)
        oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
        layer = Layer(H_CONV, (n, o, oh, ow), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w, stride=(stride, stride),
                      pad=(pad,) * 4, group=c if dw else 1)
        got = b200.run(DT_INT8, (n, c, h, w), [layer], x, s_in=0.02, zp_in=zp_in)
        want = ref.run(DT_INT8, (n, c, h, w), [layer], x, s_in=0.02, zp_in=zp_in)
        ref_band(got, want)


@pytest.mark.parametrize("case", [(2, 512, 14, 14, 512, -128), (1, 128, 28, 28, 256, 0), (3, 96, 9, 11, 200, 7),
                                  (1, 256, 20, 20, 400, 0), (2, 1024, 1, 1, 1000, 0)],
                         ids=lambda c: "n%d_c%d_%dx%d_o%d_zp%d" % c)
def test_pointwise_conv_with_cluster_multicast(case, b200, oracle, rng):
    """SHL_B200_GEMM_CLUSTER=1: CTA pairs (thread-block clusters of two) fetch each activation stage once
    and multicast it to both CTAs (csrc/gemm_tc.cu); results are the same bytes as the unpaired kernel"""
    n, c, h, w, o, zp_in = case
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, 1, 1)
    layer = Layer(H_CONV, (n, o, h, w), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w)
    os.environ["SHL_B200_GEMM_CLUSTER"] = "1"
    try:
        got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, zp_in=zp_in)
    finally:
        os.environ.pop("SHL_B200_GEMM_CLUSTER", None)
    want = oracle.conv2d_i8(x, wt, b, (n, o, h, w), stride=(1, 1), pad=(0,) * 4, dilation=(1, 1), group=1, s_in=0.02,
                            zp_in=zp_in, s_w=s_w, s_b=None, s_out=s_out, zp_out=3)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} outputs differ from the oracle"


FIRST_LAYER = [(2, 3, 32, 32, 32, 3, 2, 1, 5), (1, 3, 40, 40, 64, 7, 2, 3, 0), (1, 4, 17, 19, 24, 3, 1, 1, -9),
               (1, 1, 16, 16, 8, 5, 1, 2, 3), (3, 3, 224, 224, 32, 3, 2, 1, 0),
               (1, 3, 33, 35, 48, 3, 2, 1, -7), (5, 3, 30, 30, 16, 3, 1, 1, 11), (2, 3, 61, 47, 24, 7, 2, 3, -128),
               # row pitches that are multiples of 16 bytes: the TMA-staged tensor-core stem (others take the dp4a kernel)
               (2, 3, 48, 48, 64, 7, 2, 3, -5), (1, 3, 64, 80, 48, 7, 2, 3, 9), (2, 3, 32, 48, 16, 3, 1, 1, 11),
               (1, 3, 20, 160, 16, 3, 1, 1, 7), (1, 3, 18, 288, 32, 3, 2, 1, -3), (1, 3, 32, 32, 64, 3, 2, 1, 4),
               (3, 3, 17, 16, 8, 3, 2, 1, -128), (1, 3, 9, 272, 64, 7, 2, 3, 6)]


@pytest.mark.parametrize("direct", ["stem_tc", "dp4a", "im2col"])
@pytest.mark.parametrize("case", FIRST_LAYER, ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d_zp%d" % c)
def test_first_layer_conv_from_nchw(case, direct, b200, oracle, rng):
    """graph mode: the network input stays NCHW on the device and the first conv reads it directly,
    through the tensor-core stem kernel (csrc/conv_stem_tc.cu: 3-channel 3x3 / 7x7), the dp4a kernel
    (csrc/conv_direct.cu: every other small-K shape, or forced) or, forced, through the NCHW im2col
    gather + GEMM; a relu node with its own qinfo rides in the epilogue either way"""
    n, c, h, w, o, k, stride, pad, zp_in = case
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layers = [Layer(H_CONV, (n, o, oh, ow), s_out=s_out, zp_out=0, w=wt, b=b, s_w=s_w, stride=(stride, stride),
                    pad=(pad,) * 4), Layer(H_RELU, (n, o, oh, ow), s_out=s_out / 2, zp_out=-128)]
    if direct == "im2col":
        os.environ["SHL_B200_NO_DIRECT_CONV"] = "1"
    if direct == "dp4a":
        os.environ["SHL_B200_NO_STEM_TC"] = "1"
    try:
        got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.02, zp_in=zp_in, run_mode=RM_GRAPH)
    finally:
        os.environ.pop("SHL_B200_NO_DIRECT_CONV", None)
        os.environ.pop("SHL_B200_NO_STEM_TC", None)
    want = oracle.conv2d_i8(x, wt, b, (n, o, oh, ow), stride=(stride, stride), pad=(pad,) * 4, dilation=(1, 1),
                            group=1, s_in=0.02, zp_in=zp_in, s_w=s_w, s_b=None, s_out=s_out, zp_out=0,
                            post=(ACT_RELU, s_out / 2, -128))
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} differ"


@pytest.mark.parametrize("api", [API_RVV, API_C906, API_C920])
def test_registered_api_ids_reach_the_gpu(api, b200, oracle, rng):
    """b200 answers under the ids of the back ends it replaces (north star: thead_rvv, c9*_opt)"""
    x = rng.integers(-128, 128, size=(1, 32, 6, 6), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, 32, 32, 1, 1)
    layer = Layer(H_CONV, (1, 32, 6, 6), s_out=s_out, w=wt, b=b, s_w=s_w)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, api=api)
    want = oracle.conv2d_i8(x, wt, b, x.shape, stride=(1, 1), pad=(0,) * 4, dilation=(1, 1), group=1, s_in=0.02,
                            zp_in=0, s_w=s_w, s_b=None, s_out=s_out, zp_out=0)
    assert np.array_equal(got, want)


def test_fuse_zp2bias_and_per_tensor_weights(b200, oracle, rng):
    n, c, h, w, o, zp_in = 1, 32, 10, 10, 48, -9
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, 3, 3)
    kw = dict(stride=(1, 1), pad=(1,) * 4, dilation=(1, 1), group=1, s_in=0.02, zp_in=zp_in, s_out=s_out, zp_out=1)
    folded = (b.astype(np.int64) - zp_in * wt.astype(np.int64).sum(axis=(1, 2, 3))).astype(np.int32)
    layer = Layer(H_CONV, (n, o, h, w), s_out=s_out, zp_out=1, w=wt, b=folded, s_w=s_w, pad=(1,) * 4, fuse_zp2bias=1)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, zp_in=zp_in)
    assert np.array_equal(got, oracle.conv2d_i8(x, wt, b, (n, o, h, w), s_w=s_w, s_b=None, **kw))
    # per-tensor weight scale (quant_channel == 1) and no bias
    s1 = np.float32([1.5e-3])
    layer = Layer(H_CONV, (n, o, h, w), s_out=s_out, zp_out=1, w=wt, b=None, s_w=s1, pad=(1,) * 4)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, zp_in=zp_in)
    assert np.array_equal(got, oracle.conv2d_i8(x, wt, None, (n, o, h, w), s_w=s1, s_b=None, **kw))


@pytest.mark.parametrize("kind", [H_RELU, H_RELU6])
def test_relu_keeping_its_producers_qinfo(kind, b200, oracle, rng):
    """graph mode: a relu / relu6 node whose qinfo equals its producer's is fused as the in-domain
    clamp (no table); must equal conv -> standalone relu of the oracle, for 1x1, 3x3 and depthwise"""
    act = ACT_RELU if kind == H_RELU else ACT_RELU6
    for (c, o, k, pad, dw, zp_out) in [(64, 64, 1, 0, False, -20), (32, 48, 3, 1, False, 5), (32, 32, 3, 1, True, -128)]:
        x = rng.integers(-128, 128, size=(2, c, 12, 12), dtype=np.int8)
        wt, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k, depthwise=dw)
        if kind == H_RELU6:
            s_out = 0.11  # put 6.0 inside the int8 range so that the upper clamp matters
        layers = [Layer(H_CONV, (2, o, 12, 12), s_out=s_out, zp_out=zp_out, w=wt, b=b, s_w=s_w, pad=(pad,) * 4,
                        group=c if dw else 1), Layer(kind, (2, o, 12, 12), s_out=s_out, zp_out=zp_out)]
        got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.02, zp_in=-3, run_mode=RM_GRAPH)
        y = oracle.conv2d_i8(x, wt, b, (2, o, 12, 12), depthwise=dw, stride=(1, 1), pad=(pad,) * 4, dilation=(1, 1),
                             group=1, s_in=0.02, zp_in=-3, s_w=s_w, s_b=None, s_out=s_out, zp_out=zp_out)
        want = oracle.relu_i8(y, act, s_out, zp_out, s_out, zp_out)
        assert np.array_equal(got, want), (kind, c, o, k)


ASYM_W_CASES = [
    # n, c, h, w, o, k, stride, pad, group, depthwise, zp_in, per_channel
    (2, 64, 14, 14, 96, 1, 1, 0, 1, False, -7, True),     # pointwise GEMM, per-channel weight zero points
    (1, 512, 9, 9, 200, 1, 1, 0, 1, False, -128, True),   # 4 k-blocks, ragged N, extreme input zero point
    (2, 32, 12, 12, 48, 3, 1, 1, 1, False, 5, True),      # 3x3 with padding through im2col: padded taps cancel
    (1, 3, 32, 32, 32, 3, 2, 1, 1, False, 5, True),       # a first layer (NCHW input): im2col path, not the stem kernel
    (2, 32, 10, 10, 64, 3, 2, 1, 2, False, 0, False),     # group conv, one zero point for the whole kernel
    (1, 24, 11, 9, 24, 3, 1, 1, 1, True, -11, True),      # depthwise 3x3 (generic kernel)
    (1, 128, 28, 28, 128, 3, 2, 1, 1, True, 3, False),    # depthwise stride 2, per-tensor
]


@pytest.mark.parametrize("case", ASYM_W_CASES, ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d_g%d_dw%d_zp%d_pc%d" % c)
@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
def test_conv_int8_asymmetric_weights_bit_exact(case, mode, b200, oracle, ref_noavx, rng):
    """weight zero_point != 0 (source/nn2/utils.c:920-931; BASELINE.json configs[4] "asymmetric quant"): the GEMM
    epilogue subtracts w_zp[o] * (row sum of x~ - zp_in) per output (b200_rowsum_i8), the depthwise kernel
    accumulates the window sum per channel; bit-exact against the oracle, in band against the reference"""
    n, c, h, w, o, k, stride, pad, group, dw, zp_in, per_channel = case
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k, group=group, depthwise=dw)
    if per_channel:
        zp_w = rng.integers(-20, 21, size=o).astype(np.int32)
        zp_w[0], zp_w[-1] = -128, 127
    else:
        s_w, zp_w = s_w[:1], np.int32([9])
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layers = [Layer(H_CONV, (n, o, oh, ow), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w, zp_w=zp_w, stride=(stride, stride),
                    pad=(pad,) * 4, group=c if dw else group),
              Layer(H_RELU, (n, o, oh, ow), s_out=s_out / 2, zp_out=-128)]
    kw = dict(depthwise=dw, stride=(stride, stride), pad=(pad,) * 4, dilation=(1, 1), group=group, s_in=0.02, zp_in=zp_in,
              s_w=s_w, s_b=None, s_out=s_out, zp_out=3, zp_w=zp_w)
    got = b200.run(DT_INT8, x.shape, layers[:1], x, s_in=0.02, zp_in=zp_in, run_mode=mode)
    want = oracle.conv2d_i8(x, wt, b, (n, o, oh, ow), **kw)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} outputs differ from the oracle"
    if mode == RM_GRAPH:  # with the relu node riding in the epilogue as a table
        got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.02, zp_in=zp_in, run_mode=mode)
        want2 = oracle.conv2d_i8(x, wt, b, (n, o, oh, ow), post=(ACT_RELU, s_out / 2, -128), **kw)
        assert np.array_equal(got, want2), "fused relu table"
    else:
        ref_band(got, ref_noavx.run(DT_INT8, x.shape, layers[:1], x, s_in=0.02, zp_in=zp_in))


def test_fc_int8_asymmetric_weights_bit_exact(b200, oracle, rng):
    batch, cin, units = 37, 600, 130
    x = rng.integers(-128, 128, size=(batch, cin), dtype=np.int8)
    wt4, s_w, b, s_out = synth_conv_i8(rng, cin, units, 1, 1)
    wt = wt4.reshape(units, cin)
    zp_w = rng.integers(-30, 31, size=units).astype(np.int32)
    layer = Layer(H_FC, (batch, units), s_out=s_out, zp_out=-4, w=wt, b=b, s_w=s_w, zp_w=zp_w)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, zp_in=6)
    want = oracle.fc_i8(x, wt, b, s_in=0.02, zp_in=6, s_w=s_w, s_b=None, s_out=s_out, zp_out=-4, zp_w=zp_w)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("batch,cin,units", [(1, 1024, 1000), (8, 31, 17), (300, 2048, 100)])
def test_fc_int8_bit_exact(batch, cin, units, b200, oracle, rng):
    x = rng.integers(-128, 128, size=(batch, cin), dtype=np.int8)
    wt4, s_w, b, s_out = synth_conv_i8(rng, cin, units, 1, 1)
    wt = wt4.reshape(units, cin)
    layer = Layer(H_FC, (batch, units), s_out=s_out, zp_out=-5, w=wt, b=b, s_w=s_w)
    got = b200.run(DT_INT8, (batch, cin), [layer], x, s_in=0.02, zp_in=7)
    assert np.array_equal(got, oracle.fc_i8(x, wt, b, s_in=0.02, zp_in=7, s_w=s_w, s_b=None, s_out=s_out, zp_out=-5))


@pytest.mark.parametrize("kind,act", [(H_RELU, ACT_RELU), (H_RELU6, ACT_RELU6)])
def test_relu_int8_bit_exact(kind, act, b200, oracle, rng):
    x = rng.integers(-128, 128, size=(2, 24, 9, 11), dtype=np.int8)
    for s_in, zp_in, s_out, zp_out in [(0.037, -3, 0.0181, -128), (0.11, 5, 0.0235, -128), (0.05, 0, 0.05, 0)]:
        got = b200.run(DT_INT8, x.shape, [Layer(kind, x.shape, s_out=s_out, zp_out=zp_out)], x, s_in=s_in, zp_in=zp_in)
        assert np.array_equal(got, oracle.relu_i8(x, act, s_in, zp_in, s_out, zp_out))


def test_add_int8_bit_exact(b200, oracle, rng):
    shape = (2, 24, 9, 11)
    x = rng.integers(-128, 128, size=shape, dtype=np.int8)
    layers = [Layer(H_RELU, shape, s_out=0.021, zp_out=-128), Layer(H_ADD, shape, in0=0, in1=1, s_out=0.06, zp_out=-11)]
    r = oracle.relu_i8(x, ACT_RELU, 0.04, 3, 0.021, -128)
    want = oracle.add_i8(x, r, 0.04, 3, 0.021, -128, 0.06, -11)
    for mode in (RM_LAYER, RM_GRAPH):
        got = b200.run(DT_INT8, shape, layers, x, s_in=0.04, zp_in=3, run_mode=mode)
        assert np.array_equal(got, want), mode


POOLS = [(H_MAXPOOL, 16, 13, 15, 3, 2, 1, 0), (H_MAXPOOL, 8, 12, 12, 2, 2, 0, 0), (H_AVGPOOL, 16, 13, 15, 3, 2, 1, 0),
         (H_AVGPOOL, 8, 9, 9, 3, 1, 1, 1), (H_GAP, 40, 7, 7, 7, 1, 0, 0), (H_GAP, 1024, 7, 7, 7, 1, 0, 0)]


@pytest.mark.parametrize("case", POOLS, ids=lambda c: "op%d_c%d_%dx%d_k%d_s%d_p%d_cip%d" % c)
def test_pool_int8_bit_exact(case, b200, oracle, rng):
    kind, c, h, w, k, stride, pad, cip = case
    x = rng.integers(-128, 128, size=(2, c, h, w), dtype=np.int8)
    if kind == H_GAP:
        oh = ow = 1
        kernel, st, pd = (h, w), (1, 1), (0,) * 4
    else:
        oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
        kernel, st, pd = (k, k), (stride, stride), (pad,) * 4
    layer = Layer(kind, (2, c, oh, ow), s_out=0.043, zp_out=-20, kernel=kernel, stride=st, pad=pd, count_include_pad=cip)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.05, zp_in=9)
    want = oracle.pool_i8(x, (2, c, oh, ow), avg=kind != H_MAXPOOL, kernel=kernel, stride=st, pad=pd,
                          count_include_pad=cip, s_in=0.05, zp_in=9, s_out=0.043, zp_out=-20)
    assert np.array_equal(got, want)


def test_softmax_int8_bit_exact(b200, oracle, rng):
    x = rng.integers(-128, 128, size=(5, 1000), dtype=np.int8)
    layer = Layer(H_SOFTMAX, x.shape, s_out=1.0 / 256, zp_out=-128, axis=1)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.08, zp_in=10)
    assert np.array_equal(got, oracle.softmax_i8(x, 0.08, 10, 1.0 / 256, -128))


def test_softmax_over_channels_of_a_feature_map(b200, oracle, rng):
    """axis 1 of N x C x H x W: N * H * W rows of the pixel-major layout (int8 bit-exact in both modes, fp16 in tolerance)"""
    from test_oracle import softmax_nchw
    for shape in ((2, 21, 5, 7), (1, 150, 3, 3)):
        x = rng.integers(-128, 128, size=shape, dtype=np.int8)
        layer = Layer(H_SOFTMAX, shape, s_out=1.0 / 256, zp_out=-128, axis=1)
        want = softmax_nchw(oracle, x, 0.08, 10, 1.0 / 256, -128)
        for mode in (RM_LAYER, RM_GRAPH):
            got = b200.run(DT_INT8, shape, [layer], x, s_in=0.08, zp_in=10, run_mode=mode)
            assert np.array_equal(got, want), (shape, mode)
    xh = rng.standard_normal((2, 21, 5, 7)).astype(np.float16)
    got = b200.run(DT_F16, xh.shape, [Layer(H_SOFTMAX, xh.shape, axis=1)], xh).astype(np.float32)
    e = np.exp(xh.astype(np.float32) - xh.astype(np.float32).max(axis=1, keepdims=True))
    assert np.max(np.abs(got - e / e.sum(axis=1, keepdims=True))) < 1e-3


def test_golden_vectors_of_the_reference(golden, b200):
    """the reference's own known answers, through the GPU path: fp16 conv / dw / fc / pools / relu
    (tests/unit_test/valid_data/*.dat) and the int8 maxpool + relu vectors it ships"""
    def run(kind, xk, outk, **kw):
        x, want = golden[xk], golden[outk]
        got = b200.run(DT_F16 if x.dtype == np.float16 else DT_INT8, x.shape, [Layer(kind, want.shape, **kw)], x)
        return got, want

    for name, k, s, p in (("maxpool2x2s2", 2, 2, 0), ("maxpool3x3s2_p1", 3, 2, 1), ("maxpool3x3s1_p1", 3, 1, 1)):
        got, want = run(H_MAXPOOL, f"{name}_int8_in", f"{name}_int8_out", kernel=(k, k), stride=(s, s), pad=(p,) * 4)
        assert np.array_equal(got, want), name
    tol = 3e-2  # the fp16 goldens were accumulated in fp16 (see tests/test_oracle.py)
    g = golden
    got = b200.run(DT_F16, g["conv1x1_fp16_in"].shape, [Layer(H_CONV, (1, 19, 4, 5), w=g["conv1x1_fp16_ker"],
                                                              b=g["conv1x1_fp16_bias"])], g["conv1x1_fp16_in"])
    f16_close(got, g["conv1x1_fp16_out"], tol)
    got = b200.run(DT_F16, g["conv3x3_fp16_in"].shape, [Layer(H_CONV, (1, 19, 4, 5), w=g["conv3x3_fp16_ker"],
                                                              b=g["conv3x3_fp16_bias"], pad=(1,) * 4)], g["conv3x3_fp16_in"])
    f16_close(got, g["conv3x3_fp16_out"], tol)
    got = b200.run(DT_F16, g["dw3x3s1_fp16_in"].shape, [Layer(H_CONV, (1, 2, 4, 10), w=g["dw3x3s1_fp16_ker"],
                                                              b=g["dw3x3s1_fp16_bias"], pad=(1,) * 4, group=2)], g["dw3x3s1_fp16_in"])
    f16_close(got, g["dw3x3s1_fp16_out"], tol)
    got = b200.run(DT_F16, g["dw3x3s2_fp16_in"].shape, [Layer(H_CONV, (1, 2, 3, 9), w=g["dw3x3s2_fp16_ker"],
                                                              b=g["dw3x3s2_fp16_bias"], pad=(1,) * 4, stride=(2, 2), group=2)],
                   g["dw3x3s2_fp16_in"])
    f16_close(got, g["dw3x3s2_fp16_out"], tol)
    got = b200.run(DT_F16, (1, 17), [Layer(H_FC, (1, 31), w=g["fc_fp16_weight"], b=g["fc_fp16_bias"])], g["fc_fp16_in"])
    f16_close(got, g["fc_fp16_out"], tol)
    got, want = run(H_AVGPOOL, "avgpool2x2s2_fp16_in", "avgpool2x2s2_fp16_out", kernel=(2, 2), stride=(2, 2))
    f16_close(got, want, 2e-3)
    got, want = run(H_AVGPOOL, "avgpool3x3s2_fp16_in", "avgpool3x3s2_fp16_out", kernel=(3, 3), stride=(2, 2))
    f16_close(got, want, 2e-3)
    got, want = run(H_GAP, "global_avgpool_fp16_in", "global_avgpool_fp16_out")
    f16_close(got, want, 2e-3)
    got, want = run(H_MAXPOOL, "maxpool2x2s2_fp16_in", "maxpool2x2s2_fp16_out", kernel=(2, 2), stride=(2, 2))
    assert np.array_equal(got, want)


F16_CASES = [(1, 64, 8, 16, 64, 1, 1, 0, False), (1, 512, 9, 9, 200, 1, 1, 0, False), (1, 3, 32, 32, 32, 3, 2, 1, False),
             (2, 48, 14, 14, 56, 3, 1, 1, False), (1, 32, 14, 14, 32, 3, 1, 1, True), (1, 128, 15, 15, 128, 3, 2, 1, True),
             # register-sliding fp16 depthwise 3x3: odd widths, several row bands, ragged channels, no padding
             (3, 24, 57, 31, 24, 3, 1, 1, True), (2, 40, 112, 112, 40, 3, 2, 1, True), (1, 1024, 7, 7, 1024, 3, 1, 1, True),
             (2, 16, 9, 10, 16, 3, 1, 0, True), (1, 72, 33, 18, 72, 3, 2, 0, True), (1, 8, 1, 1, 8, 3, 1, 1, True)]


@pytest.mark.parametrize("case", F16_CASES, ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d_dw%d" % c)
def test_conv_fp16_within_tolerance(case, b200, oracle, rng):
    n, c, h, w, o, k, stride, pad, dw = case
    x = rng.standard_normal((n, c, h, w)).astype(np.float16)
    cg = 1 if dw else c
    wt = (rng.standard_normal((o, cg, k, k)) / np.sqrt(cg * k * k)).astype(np.float16)
    b = rng.standard_normal(o).astype(np.float16)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layer = Layer(H_CONV, (n, o, oh, ow), w=wt, b=b, stride=(stride, stride), pad=(pad,) * 4, group=c if dw else 1)
    got = b200.run(DT_F16, (n, c, h, w), [layer], x)
    want = oracle.conv2d_f32(x.astype(np.float32), wt.astype(np.float32), b.astype(np.float32), (n, o, oh, ow),
                             depthwise=dw, stride=(stride, stride), pad=(pad,) * 4)
    f16_close(got, want)


@pytest.mark.parametrize("case", [(2, 3, 32, 32, 32, 3, 2, 1), (1, 3, 40, 40, 64, 7, 2, 3), (3, 3, 33, 35, 48, 3, 1, 1),
                                  (1, 4, 17, 19, 24, 3, 1, 1), (2, 1, 16, 16, 8, 5, 1, 2),
                                  # 3-channel 3x3 with a row pitch of a multiple of 16 bytes: the tensor-core fp16 stem
                                  (3, 3, 224, 224, 32, 3, 2, 1), (2, 3, 24, 40, 16, 3, 1, 1), (1, 3, 18, 288, 64, 3, 2, 1),
                                  (1, 3, 20, 160, 48, 3, 1, 1), (2, 3, 17, 16, 8, 3, 2, 0)],
                         ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d" % c)
@pytest.mark.parametrize("direct", ["direct", "cuda_cores", "im2col"])
def test_first_layer_conv_fp16_from_nchw(case, direct, b200, oracle, rng):
    """graph mode, fp16: the first conv reads the NCHW graph input directly (csrc/conv_stem_f16_tc.cu: 3-channel 3x3 on
    tcgen05 kind::f16; csrc/conv_direct.cu, conv_direct_f16_kernel: f32 accumulation in registers -- every other small-K
    shape, or forced by "cuda_cores") or, forced, through im2col + GEMM; a relu node rides in the epilogue either way"""
    n, c, h, w, o, k, stride, pad = case
    x = rng.standard_normal((n, c, h, w)).astype(np.float16)
    wt = (rng.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)).astype(np.float16)
    b = rng.standard_normal(o).astype(np.float16)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layers = [Layer(H_CONV, (n, o, oh, ow), w=wt, b=b, stride=(stride, stride), pad=(pad,) * 4),
              Layer(H_RELU, (n, o, oh, ow))]
    if direct == "im2col":
        os.environ["SHL_B200_NO_DIRECT_CONV"] = "1"
    if direct == "cuda_cores":
        os.environ["SHL_B200_NO_STEM_TC"] = "1"
    try:
        got = b200.run(DT_F16, x.shape, layers, x, run_mode=RM_GRAPH)
    finally:
        os.environ.pop("SHL_B200_NO_DIRECT_CONV", None)
        os.environ.pop("SHL_B200_NO_STEM_TC", None)
    want = np.maximum(oracle.conv2d_f32(x.astype(np.float32), wt.astype(np.float32), b.astype(np.float32),
                                        (n, o, oh, ow), stride=(stride, stride), pad=(pad,) * 4), 0)
    f16_close(got, want)


def test_conv_fp16_against_the_reference_library(b200, ref, rng):
    n, c, h, w, o = 1, 64, 14, 14, 96
    x = rng.standard_normal((n, c, h, w)).astype(np.float16)
    wt = (rng.standard_normal((o, c, 3, 3)) / np.sqrt(c * 9)).astype(np.float16)
    b = rng.standard_normal(o).astype(np.float16)
    layer = Layer(H_CONV, (n, o, h, w), w=wt, b=b, pad=(1,) * 4)
    f16_close(b200.run(DT_F16, x.shape, [layer], x), ref.run(DT_F16, x.shape, [layer], x))


@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
def test_mobilenet_v1_int8_narrow_bit_exact(mode, b200):
    """a whole (narrow, 64x64) MobileNetV1: conv3x3 s2 from NCHW, 13 dw + 13 pw with relu fused in
    graph mode, global avgpool, 1x1 classifier, softmax; batch 3 so that image boundaries matter"""
    nb = nets.mobilenet_v1(DT_INT8, batch=3, res=64, width=0.25, classes=100)
    x = nb.input_batch()
    got = b200.run(DT_INT8, nb.in_shape, nb.layers, x, s_in=nb.s_in, zp_in=nb.zp_in, run_mode=mode)
    want = nets.oracle_forward(nb, x)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} differ"


def test_mobilenet_v1_int8_full_size_graph(b200, ref):
    """BASELINE.json configs[1]: MobileNetV1 int8, NCHW batch 1, example graph shapes -- bit-exact
    against the oracle chain, and within the reference's noise band against the reference itself
    run through GREF (source/graph_ref/setup.c:1305) on the same tensors"""
    nb = nets.mobilenet_v1(DT_INT8, batch=1)
    x = nb.input_batch()
    with b200.create(DT_INT8, nb.in_shape, nb.layers, s_in=nb.s_in, zp_in=nb.zp_in, run_mode=RM_GRAPH, api=API_C906) as net:
        got = net(x)
        again = net(x)
    want = nets.oracle_forward(nb, x)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} differ"
    assert np.array_equal(got, again), "replaying the CUDA graph is not idempotent"
    got_ref = ref.run(DT_INT8, nb.in_shape, nb.layers, x, s_in=nb.s_in, zp_in=nb.zp_in, run_mode=RM_GRAPH)
    d = np.abs(got.astype(int) - got_ref.astype(int))
    assert d.max() <= 1 and np.count_nonzero(d) <= 5, (d.max(), np.count_nonzero(d))
    assert len(np.unique(got)) > 3, "degenerate output"


def test_mobilenet_v1_int8_batch_properties(b200):
    """BASELINE.json's headline size (batch 256) is too slow for the scalar oracle, so it is checked
    through properties: every image of a batch equals the same image run alone (images are
    independent units -- the property batch sharding rests on), replay is idempotent, and one image
    is checked against the oracle chain."""
    batch = 256
    nb = nets.mobilenet_v1(DT_INT8, batch=batch)
    x = nb.input_batch()
    x[7] = x[200]  # two equal images inside the batch must give equal outputs
    with b200.create(DT_INT8, nb.in_shape, nb.layers, s_in=nb.s_in, zp_in=nb.zp_in, run_mode=RM_GRAPH) as net:
        y = net(x)
        y2 = net(x)
    assert np.array_equal(y, y2)
    assert np.array_equal(y[7], y[200])
    nb1 = nets.mobilenet_v1(DT_INT8, batch=1)
    with b200.create(DT_INT8, nb1.in_shape, nb1.layers, s_in=nb1.s_in, zp_in=nb1.zp_in, run_mode=RM_GRAPH) as net1:
        for i in (0, 7, 255):
            assert np.array_equal(net1(x[i:i + 1])[0], y[i]), f"image {i}: batched != alone"
    assert np.array_equal(y[3:4], nets.oracle_forward(nb1, x[3:4]))


@pytest.mark.parametrize("first_op", ["conv", "dw"])
def test_prefetched_inputs_give_the_same_bytes(first_op, b200, rng):
    """shl_b200_session_prefetch_input: batch k+1 travels on the copy stream while batch k computes;
    results must equal the plain csinn_update_input + csinn_session_run sequence, for an input that
    stays NCHW on the device (first op a dense conv) and one converted to pixel-major per run, and
    a pointer that was prefetched but is NOT the next input must not be picked up"""
    if first_op == "conv":
        nb = nets.mobilenet_v1(DT_INT8, batch=4, res=64, width=0.25, classes=50)
        shape, layers, s_in, zp_in = nb.in_shape, nb.layers, nb.s_in, nb.zp_in
    else:
        c = 48
        wt, s_w, b, s_out = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
        shape, s_in, zp_in = (4, c, 20, 20), 0.02, -3
        layers = [Layer(H_CONV, shape, s_out=s_out, zp_out=5, w=wt, b=b, s_w=s_w, stride=(1, 1), pad=(1,) * 4, group=c)]
    xs = [rng.integers(-128, 128, size=shape, dtype=np.int8) for _ in range(5)]
    with b200.create(DT_INT8, shape, layers, s_in=s_in, zp_in=zp_in, run_mode=RM_GRAPH) as net:
        plain = [net(x) for x in xs]
        piped = net.stream(xs, prefetch=True)
        for k in range(5):
            assert np.array_equal(plain[k], piped[k]), f"batch {k} differs through the prefetch stage"
        # a stale prefetch: stage xs[0], then run xs[1] -- must compute xs[1]
        L = b200.lib
        a, bb = np.ascontiguousarray(xs[0]), np.ascontiguousarray(xs[1])
        assert L.h_net_prefetch_input(net.handle, a.ctypes.data) == 0
        assert np.array_equal(net.stream([bb], prefetch=False)[0], plain[1])
        assert np.array_equal(net.stream([a], prefetch=False)[0], plain[0])


def test_mobilenet_v1_fp16_graph(b200):
    """BASELINE.json configs[2] shapes (c906_mobilenetv1_f16.c), fp16, at a size the oracle finishes: the class
    scores BEFORE the softmax (probabilities of ~5e-3 hide everything behind an absolute tolerance), then the
    probabilities.  28 layers each rounded to fp16 (the oracle chain rounds per layer too, but sums in another
    order) accumulate, so the end-to-end bound on the scores is 4e-3; every layer on its own is held to the
    north-star 1e-3 in test_mobilenet_v1_fp16_full_size_every_layer."""
    nb = nets.mobilenet_v1(DT_F16, batch=2, res=96, width=0.5, classes=200, softmax=False)
    x = nb.input_batch()
    got = b200.run(DT_F16, nb.in_shape, nb.layers, x, run_mode=RM_GRAPH, api=API_C906)
    want = nets.oracle_forward(nb, x)
    f16_close(got, want, 4e-3)
    assert float(np.abs(want.astype(np.float32)).max()) > 0.5, "degenerate scores"
    nb = nets.mobilenet_v1(DT_F16, batch=2, res=96, width=0.5, classes=200)
    got = b200.run(DT_F16, nb.in_shape, nb.layers, x, run_mode=RM_GRAPH, api=API_C906)
    want = nets.oracle_forward(nb, x)
    assert np.max(np.abs(got.astype(np.float32) - want.astype(np.float32))) < 1e-3
    assert abs(float(got.astype(np.float32).sum()) - 2.0) < 2e-2


def test_mobilenet_v1_fp16_full_size_every_layer(b200):
    """BASELINE.json configs[2] at its true layer shapes (224 x 224, width 1): every layer of the graph is run
    on the device from the ORACLE's input for that layer and held to 1e-3 relative against f32 math -- the
    north-star tolerance, per operator, at full size"""
    nb = nets.mobilenet_v1(DT_F16, batch=1, softmax=False)
    x = nb.input_batch()
    vals = nets.oracle_forward(nb, x, all_values=True)
    worst = 0.0
    for i, l in enumerate(nb.layers):
        a = vals[l.in0 if l.in0 >= 0 else i]
        lone = Layer(**{**l.__dict__, "in0": -1})
        got = b200.run(DT_F16, a.shape, [lone], a)
        g, w = got.astype(np.float32), vals[i + 1].astype(np.float32)
        err = float((np.abs(g - w) / np.maximum(np.abs(w), 1.0)).max())
        worst = max(worst, err)
        assert err <= F16_TOL, f"layer {i} (kind {l.kind}, out {l.out_shape}): max relative error {err:.3e}"
    print(f"fp16 MobileNetV1 full size: worst per-layer relative error {worst:.2e}")


def test_mobilenet_v1_fp16_batch_256_properties(b200):
    """BASELINE.json configs[2] at its batch: every image of the batch equals the same image run alone (bytes),
    replay is idempotent, and one image's class scores are within the end-to-end bound of the oracle chain"""
    batch = 256
    nb = nets.mobilenet_v1(DT_F16, batch=batch, softmax=False)
    x = nb.input_batch()
    x[9] = x[130]
    with b200.create(DT_F16, nb.in_shape, nb.layers, run_mode=RM_GRAPH) as net:
        y = net(x)
        assert np.array_equal(y.view(np.uint16), net(x).view(np.uint16))
    assert np.array_equal(y[9].view(np.uint16), y[130].view(np.uint16))
    nb1 = nets.mobilenet_v1(DT_F16, batch=1, softmax=False)
    with b200.create(DT_F16, nb1.in_shape, nb1.layers, run_mode=RM_GRAPH) as net1:
        for i in (0, 9, 255):
            assert np.array_equal(net1(x[i:i + 1])[0].view(np.uint16), y[i].view(np.uint16)), f"image {i}: batched != alone"
    f16_close(y[3:4], nets.oracle_forward(nb1, x[3:4]), 4e-3)


def test_resnet50_int8_full_size_bit_exact(b200):
    """BASELINE.json configs[4] at its true layer shapes (224 x 224, width 1, 1000 classes), batch 2: all 16
    bottlenecks, every 3x3 / strided / residual shape, bit-exact against the oracle chain"""
    nb = nets.resnet50(DT_INT8, batch=2)
    x = nb.input_batch()
    got = b200.run(DT_INT8, nb.in_shape, nb.layers, x, s_in=nb.s_in, zp_in=nb.zp_in, run_mode=RM_GRAPH)
    want = nets.oracle_forward(nb, x)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} differ"
    assert len(np.unique(got)) > 10, "degenerate output"


def test_resnet50_int8_narrow_bit_exact(b200):
    """BASELINE.json configs[4] structure (7x7 s2 stem, maxpool, bottlenecks with strided 1x1
    shortcuts, add + relu, global avgpool, flatten, fullyconnected), narrow and small"""
    nb = nets.resnet50(DT_INT8, batch=2, res=64, width=0.25, classes=50)
    x = nb.input_batch()
    got = b200.run(DT_INT8, nb.in_shape, nb.layers, x, s_in=nb.s_in, zp_in=nb.zp_in, run_mode=RM_GRAPH)
    want = nets.oracle_forward(nb, x)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} differ"


@pytest.mark.parametrize("rows", ["imma", "igemm", "umma128", "umma", "tma", "generic"])
def test_dwconv3x3_kernel_variants(rows, b200, oracle, rng):
    """the tensor-core depthwise kernel (csrc/dwconv3x3_umma.cu: stride 1, "same" padding; other cases fall
    through), the TMA-fed dp4a kernel (csrc/dwconv3x3_tma.cu) and the generic one (csrc/dwconv.cu)
    on shapes that hit ragged rows / columns / channel chunks, both strides, zero-point patching of
    the halo, unpadded borders and a fused relu table"""
    # "imma" (the default path): taps as diagonal-B warp MMAs (csrc/dwconv3x3_imma.cu) where channels are a multiple of
    # 32 and the map is at least 12 wide; everything else falls through to the dp4a kernel, which "tma" forces
    os.environ["SHL_B200_DW_IMMA"] = "2" if rows == "imma" else "0"  # 2: wherever it applies, not only where it is faster
    # "igemm": the implicit-GEMM kernel against tap-diagonal weights where channels are a multiple of 64 (opt-in)
    if rows == "igemm":
        os.environ["SHL_B200_DW_IGEMM"] = "1"
    if rows == "generic":
        os.environ["SHL_B200_DW_GENERIC"] = "1"
    if rows == "umma":
        os.environ["SHL_B200_DW_UMMA"] = "1"
    if rows == "umma128":  # csrc/dwconv3x3_umma128.cu: channel counts that are multiples of 128
        os.environ["SHL_B200_DW_UMMA"] = "2"
    try:
        for (n, c, h, w, stride, pad, zp_in) in [(2, 32, 13, 29, 1, 1, -7), (7, 48, 7, 7, 1, 1, 5), (5, 16, 3, 2, 1, 1, -3),
                                                 (1, 40, 1, 1, 1, 1, 9), (1, 16, 40, 200, 1, 1, -9), (2, 128, 13, 29, 1, 1, -7),
                                                 (5, 256, 7, 7, 1, 1, 3), (2, 128, 56, 56, 1, 1, -128), (1, 1024, 7, 7, 1, 1, 9),
                                                 (300, 128, 14, 14, 1, 1, -6), (1, 64, 56, 56, 1, 1, 0),
                                                 (1, 48, 15, 15, 2, 1, 4), (1, 16, 7, 7, 1, 1, -128),
                                                 (1, 128, 9, 10, 2, 0, 2), (1, 20, 6, 5, 1, 0, 1),
                                                 (2, 96, 30, 33, 1, 1, -5), (1, 160, 57, 9, 2, 1, 7),
                                                 (1, 512, 14, 14, 1, 1, -128), (3, 32, 112, 112, 1, 1, -128),
                                                 (2, 64, 30, 33, 2, 1, -9), (1, 128, 57, 40, 2, 1, 11), (2, 32, 28, 28, 2, 1, -128),
                                                 (1, 256, 28, 28, 1, 1, 6), (1, 64, 112, 112, 2, 1, -128), (1, 128, 9, 30, 1, 0, -4),
                                                 (2, 256, 29, 31, 2, 0, 8), (1, 64, 17, 12, 1, 1, -2), (3, 512, 14, 14, 2, 1, 5),
                                                 (1, 192, 20, 45, 1, 1, -11), (1, 96, 33, 18, 2, 1, 3)]:
            x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
            wt, s_w, b, s_out = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
            oh, ow = conv_out_hw(h, w, 3, 3, (stride, stride), (pad,) * 4)
            layers = [Layer(H_CONV, (n, c, oh, ow), s_out=s_out, zp_out=2, w=wt, b=b, s_w=s_w, stride=(stride, stride),
                            pad=(pad,) * 4, group=c), Layer(H_RELU, (n, c, oh, ow), s_out=s_out / 2, zp_out=-128)]
            kw = dict(depthwise=True, stride=(stride, stride), pad=(pad,) * 4, dilation=(1, 1), group=1, s_in=0.02,
                      zp_in=zp_in, s_w=s_w, s_b=None, s_out=s_out, zp_out=2)
            got = b200.run(DT_INT8, x.shape, layers[:1], x, s_in=0.02, zp_in=zp_in)
            assert np.array_equal(got, oracle.conv2d_i8(x, wt, b, (n, c, oh, ow), **kw)), (rows, n, c, h, w, stride)
            got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.02, zp_in=zp_in, run_mode=RM_GRAPH)
            want = oracle.conv2d_i8(x, wt, b, (n, c, oh, ow), post=(ACT_RELU, s_out / 2, -128), **kw)
            assert np.array_equal(got, want), (rows, n, c, h, w, stride, "fused relu")
    finally:
        os.environ.pop("SHL_B200_DW_GENERIC", None)
        os.environ.pop("SHL_B200_DW_UMMA", None)
        os.environ.pop("SHL_B200_DW_IGEMM", None)
        os.environ.pop("SHL_B200_DW_IMMA", None)


def test_dwconv_sweep_shapes(b200, oracle, rng):
    """BASELINE.json configs[3] (depthwise 3x3 int8, H=W=56) at batch 1 per channel count"""
    for c in (32, 64, 128, 256, 512, 1024):
        x = rng.integers(-128, 128, size=(1, c, 56, 56), dtype=np.int8)
        wt, s_w, b, s_out = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
        layer = Layer(H_CONV, (1, c, 56, 56), s_out=s_out, zp_out=-4, w=wt, b=b, s_w=s_w, pad=(1,) * 4, group=c)
        got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, zp_in=-7)
        want = oracle.conv2d_i8(x, wt, b, x.shape, depthwise=True, stride=(1, 1), pad=(1,) * 4, dilation=(1, 1),
                                group=1, s_in=0.02, zp_in=-7, s_w=s_w, s_b=None, s_out=s_out, zp_out=-4)
        assert np.array_equal(got, want), c


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["softmax_like", "few_maxima", "wide_range", "ties", "tiny"])
def test_softmax_denominator_on_the_device_equals_the_literal_loop(mode):
    """the softmax kernel's denominator (parallel integer scan per binade of the running sum,
    csrc/softmax.cu + softmax_sum.h) against the reference's literal `float acc += double` loop
    (source/reference/softmax.c:53-55) evaluated in numpy: the float sums must have the same bits --
    the int8 / fp16 outputs of the op are too coarse to show a wrong last bit of the denominator"""
    import ctypes as C
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = C.CDLL(os.path.join(root, "csi-nn2_b200", "lib", "libb200nn.so"))
    shim.b200_last_error.restype = C.c_char_p
    rng = np.random.default_rng({"softmax_like": 1, "few_maxima": 2, "wide_range": 3, "ties": 4, "tiny": 5}[mode])
    rows = 48
    for c in (1, 7, 255, 256, 257, 1000, 2048):
        u = rng.random((rows, c))
        if mode == "softmax_like":
            e = np.exp(-20.0 * u)
        elif mode == "few_maxima":
            e = np.where(rng.random((rows, c)) < 0.02, 1.0, np.exp(-40.0 * u))
        elif mode == "wide_range":
            e = np.ldexp(u, -(60.0 * rng.random((rows, c))).astype(np.int32))
        elif mode == "ties":
            e = np.ldexp((1 + (7 * u).astype(np.int64)).astype(np.float64), -24 - (3 * rng.random((rows, c))).astype(np.int32))
        else:
            e = np.where(rng.random((rows, c)) < 0.5, 0.0, np.ldexp(u, -140))
        e = np.ascontiguousarray(e, dtype=np.float64)
        want = np.zeros(rows, np.float32)
        for r in range(rows):
            acc = np.float32(0)
            for v in e[r]:
                acc = np.float32(np.float64(acc) + v)
            want[r] = acc
        d_e, d_o = C.c_void_p(), C.c_void_p()
        assert shim.b200_malloc(C.byref(d_e), C.c_size_t(e.nbytes)) == 0, shim.b200_last_error()
        assert shim.b200_malloc(C.byref(d_o), C.c_size_t(4 * rows)) == 0, shim.b200_last_error()
        got = np.zeros(rows, np.float32)
        assert shim.b200_memcpy_h2d(d_e, e.ctypes.data_as(C.c_void_p), C.c_size_t(e.nbytes), None) == 0
        assert shim.b200_test_softmax_denominator(d_e, rows, c, d_o, None) == 0, shim.b200_last_error()
        assert shim.b200_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_o, C.c_size_t(4 * rows), None) == 0
        assert shim.b200_stream_sync(None) == 0, shim.b200_last_error()
        shim.b200_free(d_e), shim.b200_free(d_o)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (mode, c, int(np.count_nonzero(got != want)))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [DT_INT8, DT_F16], ids=["int8", "fp16"])
def test_binary_model_saved_and_restored_on_b200(dtype, b200, tmp_path):
    """HHB binary model format: a graph-mode MobileNetV1 (narrow, 64x64) built
    through the API under the C920 id with save_mode = CSINN_SAVE_AND_RUN writes itself at
    session_setup (b200_opt/graph.c, the reference's own serialisers); csinn_import_binary_model +
    csinn_load_binary_model (CSINN_LOAD_BG -> shl_b200_load_binary_model) restore it into a fresh
    session whose outputs are the same bytes; the restored int8 model also equals the oracle chain"""
    nb = nets.mobilenet_v1(dtype, batch=2, res=64, width=0.25, classes=40)
    x = nb.input_batch()
    path = str(tmp_path / "b200_model.bm")
    b200.save_next(path)
    net = b200.create(dtype, nb.in_shape, nb.layers, s_in=nb.s_in, zp_in=nb.zp_in, run_mode=RM_GRAPH, api=API_C920)
    try:
        want = net(x)
    finally:
        net.close()
    blob = open(path, "rb").read()
    assert len(blob) > 8192
    restored = b200.import_model(blob, dtype, nb.in_shape, want.shape)
    try:
        got = restored(x)
        again = restored(x)
    finally:
        restored.close()
    assert np.array_equal(got, want) and np.array_equal(again, want)
    if dtype == DT_INT8:
        assert np.array_equal(got.reshape(2, -1), nets.oracle_forward(nb, x).reshape(2, -1))


UNARY = [("leaky", 14, 3, 0.1, 0.0), ("sigmoid", 15, 4, 0.0, 0.0), ("clip", 16, 5, -0.5, 1.25), ("silu", 20, 6, 0.0, 0.0),
         ("erf", 21, 7, 0.0, 0.0)]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
@pytest.mark.parametrize("name,kind,op,p0,p1", UNARY, ids=[u[0] for u in UNARY])
def test_unary_ops_int8_bit_exact(name, kind, op, p0, p1, mode, b200, oracle, rng):
    """leaky relu / sigmoid / clip, int8: standalone (a 256-entry table of the reference's float
    sequence) and, in graph mode, riding in the producing convolution's epilogue"""
    x = rng.integers(-128, 128, size=(2, 24, 9, 11), dtype=np.int8)
    got = b200.run(DT_INT8, x.shape, [Layer(kind, x.shape, s_out=0.043, zp_out=-20, p0=p0, p1=p1)], x, s_in=0.05,
                   zp_in=9, run_mode=mode)
    assert np.array_equal(got, oracle.unary_i8(x, op, p0, p1, 0.05, 9, 0.043, -20))
    # conv -> unary (fused in graph mode, two kernels in layer mode): same bytes as oracle conv then oracle unary
    n, c, h, w, o = 2, 32, 10, 10, 48
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, 1, 1)
    layers = [Layer(H_CONV, (n, o, h, w), s_out=s_out, zp_out=4, w=wt, b=b, s_w=s_w),
              Layer(kind, (n, o, h, w), s_out=s_out / 3, zp_out=-100, p0=p0, p1=p1)]
    got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.02, zp_in=-5, run_mode=mode)
    mid = oracle.conv2d_i8(x, wt, b, (n, o, h, w), stride=(1, 1), pad=(0,) * 4, dilation=(1, 1), group=1, s_in=0.02,
                           zp_in=-5, s_w=s_w, s_b=None, s_out=s_out, zp_out=4)
    assert np.array_equal(got, oracle.unary_i8(mid, op, p0, p1, s_out, 4, s_out / 3, -100))


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind,op,p0,p1", UNARY, ids=[u[0] for u in UNARY])
def test_unary_ops_fp16_within_tolerance(name, kind, op, p0, p1, b200, rng):
    x = (3.0 * rng.standard_normal((2, 24, 9, 11))).astype(np.float16)
    got = b200.run(DT_F16, x.shape, [Layer(kind, x.shape, p0=p0, p1=p1)], x).astype(np.float32)
    xf = x.astype(np.float32)
    want = {3: np.where(xf > 0, xf, xf * np.float32(p0)), 4: 1.0 / (1.0 + np.exp(-xf.astype(np.float64))),
            5: np.clip(xf, p0, p1), 6: xf.astype(np.float64) / (1.0 + np.exp(-xf.astype(np.float64))),
            7: np.vectorize(__import__("math").erf)(xf.astype(np.float64))}[op].astype(np.float32)
    f16_close(got.astype(np.float16), want)


@pytest.mark.gpu
def test_graph_mode_fuses_unary_nodes_into_their_producer(b200, rng):
    """graph.c: conv -> relu / leaky relu / sigmoid / clip with the unary node as the only consumer
    becomes ONE step (int8: the node is the epilogue's post table); the session reports its steps"""
    import ctypes as C
    shl = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csi-nn2_b200", "lib",
                              "libshl_b200.so"))
    shl.shl_b200_session_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    n, c, h, w, o = 1, 32, 8, 8, 32
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, 1, 1)
    for kind, p0, p1 in [(H_RELU, 0, 0), (14, 0.1, 0), (15, 0, 0), (16, -0.5, 1.0), (20, 0, 0), (21, 0, 0)]:
        layers = [Layer(H_CONV, (n, o, h, w), s_out=s_out, zp_out=4, w=wt, b=b, s_w=s_w),
                  Layer(kind, (n, o, h, w), s_out=s_out / 3, zp_out=-100, p0=p0, p1=p1)]
        net = b200.create(DT_INT8, (n, c, h, w), layers, s_in=0.02, zp_in=-5, run_mode=RM_GRAPH)
        try:
            buf = C.create_string_buffer(4096)
            shl.shl_b200_session_describe(net.session, buf, len(buf))
            head = buf.value.decode().splitlines()[0]
            assert "steps=1 " in head, head
        finally:
            net.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,op,s_out,zp_out", [(17, 1, 0.05, -11), (18, 2, 0.012, -30), (25, 4, 0.09, 4)],
                         ids=["sub", "mul", "div"])
def test_sub_mul_int8_bit_exact(kind, op, s_out, zp_out, b200, oracle, rng):
    shape = (2, 24, 9, 11)
    x = rng.integers(-128, 128, size=shape, dtype=np.int8)
    layers = [Layer(H_RELU, shape, s_out=0.021, zp_out=-128), Layer(kind, shape, in0=0, in1=1, s_out=s_out, zp_out=zp_out)]
    r = oracle.relu_i8(x, ACT_RELU, 0.04, 3, 0.021, -128)
    want = oracle.binary_i8(op, x, r, 0.04, 3, 0.021, -128, s_out, zp_out)
    for mode in (RM_LAYER, RM_GRAPH):
        got = b200.run(DT_INT8, shape, layers, x, s_in=0.04, zp_in=3, run_mode=mode)
        assert np.array_equal(got, want), mode


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [17, 18], ids=["sub", "mul"])
def test_sub_mul_fp16_within_tolerance(kind, b200, rng):
    shape = (2, 24, 9, 11)
    x = rng.standard_normal(shape).astype(np.float16)
    layers = [Layer(H_RELU, shape), Layer(kind, shape, in0=0, in1=1)]
    got = b200.run(DT_F16, shape, layers, x, run_mode=RM_GRAPH)
    xf = x.astype(np.float32)
    r = np.maximum(xf, 0)
    f16_close(got, xf - r if kind == 17 else xf * r)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,op", [(H_ADD, 0), (17, 1), (18, 2)], ids=["add", "sub", "mul"])
def test_binary_ops_on_rounding_ties_and_saturation(kind, op, b200, oracle, rng):
    """the int8 binary kernel decides rint() of the IEEE quotient from r * RN(1 / s_out) and takes the
    real division only near half-integers (csrc/eltwise.cu): scales chosen so that the quotient IS a
    half-integer (or within an ulp of one) for about half of all input pairs, plus scales that saturate"""
    shape = (4, 32, 24, 24)
    x = rng.integers(-128, 128, size=shape, dtype=np.int8)
    for s_a, zp_a, s_r, zp_r, s_out, zp_out in [(0.02, 0, 0.02, -128, 0.04, 0), (0.03, 3, 0.015, -128, 0.06, -7),
                                                (0.5, -1, 0.25, -128, 0.125, 5), (0.001, 0, 0.003, -128, 0.5, 0),
                                                (1.0, 0, 1.0, -128, 1.0 / 3.0, 0)]:
        layers = [Layer(H_RELU, shape, s_out=s_r, zp_out=zp_r), Layer(kind, shape, in0=0, in1=1, s_out=s_out, zp_out=zp_out)]
        r = oracle.relu_i8(x, ACT_RELU, s_a, zp_a, s_r, zp_r)
        want = oracle.binary_i8(op, x, r, s_a, zp_a, s_r, zp_r, s_out, zp_out)
        got = b200.run(DT_INT8, shape, layers, x, s_in=s_a, zp_in=zp_a, run_mode=RM_GRAPH)
        assert np.array_equal(got, want), (s_a, s_r, s_out, int(np.count_nonzero(got != want)))
        # the same op with a relu behind it: the graph planner folds the relu into the binary kernel as a 256-entry
        # table (for add / sub this is the specialised residual-add kernel's table variant)
        layers2 = layers + [Layer(H_RELU, shape, s_out=s_out * 0.7, zp_out=-128)]
        want2 = oracle.relu_i8(want, ACT_RELU, s_out, zp_out, s_out * 0.7, -128)
        got2 = b200.run(DT_INT8, shape, layers2, x, s_in=s_a, zp_in=zp_a, run_mode=RM_GRAPH)
        assert np.array_equal(got2, want2), ("relu", s_a, s_r, s_out, int(np.count_nonzero(got2 != want2)))


from test_oracle import CONCAT_CASES, concat_case


@pytest.mark.gpu
@pytest.mark.parametrize("shape,axis,three", CONCAT_CASES)
def test_concat_int8_bit_exact(shape, axis, three, b200, oracle, rng):
    x, layers, want = concat_case(shape, axis, three, oracle, rng)
    for mode in (RM_LAYER, RM_GRAPH):
        got = b200.run(DT_INT8, shape, layers, x, s_in=0.04, zp_in=3, run_mode=mode)
        assert np.array_equal(got, want), mode


@pytest.mark.gpu
@pytest.mark.parametrize("shape,axis,three", CONCAT_CASES)
def test_concat_fp16_is_a_copy(shape, axis, three, b200, rng):
    from shl import H_CONCAT
    x = rng.standard_normal(shape).astype(np.float16)
    out_shape = list(shape)
    out_shape[axis] *= 3 if three else 2
    layers = [Layer(H_RELU, shape), Layer(H_CONCAT, tuple(out_shape), in0=0, in1=1, axis=axis, p0=3.0 if three else 0.0)]
    r = np.maximum(x, np.float16(0))
    want = np.concatenate([x, r, x] if three else [x, r], axis=axis)
    for mode in (RM_LAYER, RM_GRAPH):
        got = b200.run(DT_F16, shape, layers, x, run_mode=mode)
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), mode


@pytest.mark.gpu
def test_concat_feeds_a_convolution_in_graph_mode(b200, oracle, rng):
    """an inception-style join: the channel padding lanes of the concatenated tensor are read by the next GEMM"""
    from shl import H_CONCAT
    shape = (2, 13, 6, 6)
    x = rng.integers(-128, 128, size=shape, dtype=np.int8)
    o, c = 24, 26
    w = rng.integers(-127, 128, size=(o, c, 1, 1), dtype=np.int8)
    s_w = (1e-3 * (1 + np.arange(o) / o)).astype(np.float32)
    b = rng.integers(-1000, 1000, size=o, dtype=np.int32)
    layers = [Layer(H_RELU, shape, s_out=0.021, zp_out=-128),
              Layer(H_CONCAT, (2, c, 6, 6), in0=0, in1=1, s_out=0.033, zp_out=5, axis=1),
              Layer(H_CONV, (2, o, 6, 6), w=w, b=b, s_w=s_w, s_out=0.05, zp_out=-3)]
    r = oracle.relu_i8(x, ACT_RELU, 0.04, 3, 0.021, -128)
    cat = oracle.concat_i8([x, r], [(0.04, 3), (0.021, -128)], 1, 0.033, 5)
    want = oracle.conv2d_i8(cat, w, b, (2, o, 6, 6), stride=(1, 1), pad=(0,) * 4, dilation=(1, 1), group=1, s_in=0.033,
                            zp_in=5, s_w=s_w, s_b=None, s_out=0.05, zp_out=-3)
    got = b200.run(DT_INT8, shape, layers, x, s_in=0.04, zp_in=3, run_mode=RM_GRAPH)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_global_maxpool_int8_bit_exact_and_fp16(b200, oracle, rng):
    from shl import H_GMP
    for shape in [(2, 24, 7, 7), (1, 5, 13, 3), (3, 1024, 7, 7)]:
        out_shape = (shape[0], shape[1], 1, 1)
        x = rng.integers(-128, 128, size=shape, dtype=np.int8)
        want = oracle.pool_i8(x, out_shape, avg=False, kernel=shape[2:], stride=(1, 1), pad=(0, 0), count_include_pad=0,
                              s_in=0.04, zp_in=3, s_out=0.03, zp_out=-9)
        for mode in (RM_LAYER, RM_GRAPH):
            got = b200.run(DT_INT8, shape, [Layer(H_GMP, out_shape, s_out=0.03, zp_out=-9)], x, s_in=0.04, zp_in=3,
                           run_mode=mode)
            assert np.array_equal(got, want), (shape, mode)
        xh = rng.standard_normal(shape).astype(np.float16)
        got = b200.run(DT_F16, shape, [Layer(H_GMP, out_shape)], xh, run_mode=RM_GRAPH)
        assert np.array_equal(got.reshape(shape[0], shape[1]), xh.max(axis=(2, 3)))


from test_oracle import BCAST_CASES, bcast_case


@pytest.mark.gpu
@pytest.mark.parametrize("kind,op,scalar", BCAST_CASES)
def test_binary_ops_with_a_constant_operand_int8_bit_exact(kind, op, scalar, b200, oracle, rng):
    for shape in [(2, 24, 5, 7), (1, 64, 9, 9)]:
        x, layer, want = bcast_case(kind, op, scalar, oracle, rng, shape)
        for mode in (RM_LAYER, RM_GRAPH):
            got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.04, zp_in=3, run_mode=mode)
            assert np.array_equal(got, want), (shape, mode)


@pytest.mark.gpu
def test_binary_ops_with_a_constant_operand_fp16(b200, rng):
    shape = (2, 24, 5, 7)
    x = rng.standard_normal(shape).astype(np.float16)
    k = rng.standard_normal(shape[1]).astype(np.float16)
    for kind, fn in [(H_ADD, np.add), (17, np.subtract), (18, np.multiply), (23, lambda a, s: np.where(a >= 0, a, a * s))]:
        got = b200.run(DT_F16, shape, [Layer(kind, shape, in0=0, w=k)], x, run_mode=RM_GRAPH)
        f16_close(got, fn(x.astype(np.float32), k.astype(np.float32).reshape(1, -1, 1, 1)))


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(1, 64, 8, 16, 64, 1, 1, 0, False), (2, 48, 14, 14, 56, 3, 1, 1, False),
                                  (1, 32, 14, 14, 32, 3, 1, 1, True), (1, 3, 32, 32, 32, 3, 2, 1, False),
                                  (1, 256, 14, 14, 200, 1, 1, 0, False)],
                         ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_p%d_dw%d" % c)
def test_conv_fp16_activations_int8_weights(case, b200, ref, oracle, rng):
    """CSINN_QUANT_FLOAT16_W_INT8 (c906_opt/fp16/convolution.c:77-81): fp16 tensors over int8
    per-channel weights, dequantised to fp16 at init -- against the f32 oracle on the dequantised weights and against
    the reference library fed the same mixed-dtype tensors.  On the tensor-core GEMM path (everything but depthwise
    and first-layer shapes) the weights stay EXACT integers in fp16 and the per-channel scale is applied in the f32
    epilogue, so the north-star 1e-3 holds; depthwise / first-layer kernels still take weights rounded to fp16
    (2^-11 relative each, as on the C906) and are held to 2e-3"""
    n, c, h, w, o, k, stride, pad, dw = case
    x = rng.standard_normal((n, c, h, w)).astype(np.float16)
    cg = 1 if dw else c
    wq = rng.integers(-127, 128, size=(o, cg, k, k), dtype=np.int8)
    s_w = ((1.0 + np.arange(o) / o) / (127.0 * np.sqrt(cg * k * k))).astype(np.float32)
    b = rng.standard_normal(o).astype(np.float16)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layer = Layer(H_CONV, (n, o, oh, ow), w=wq, b=b, s_w=s_w, stride=(stride, stride), pad=(pad,) * 4, group=c if dw else 1)
    for mode in (RM_LAYER, RM_GRAPH):
        got = b200.run(DT_F16, (n, c, h, w), [layer], x, run_mode=mode)
        wf = wq.astype(np.float32) * s_w.reshape(-1, 1, 1, 1)
        want = oracle.conv2d_f32(x.astype(np.float32), wf, b.astype(np.float32), (n, o, oh, ow), depthwise=dw,
                                 stride=(stride, stride), pad=(pad,) * 4)
        exact_weights = not dw and not (cg * k * k <= 160 and o <= 64)
        f16_close(got, want, tol=F16_TOL if exact_weights else 2e-3)
    if not dw and n == 1:  # the AVX reference path reads batch 1 only
        f16_close(got, ref.run(DT_F16, (n, c, h, w), [layer], x), tol=F16_TOL if exact_weights else 2e-3)


@pytest.mark.gpu
def test_fullyconnected_fp16_activations_int8_weights(b200, rng):
    n, d, o = 3, 200, 75
    x = rng.standard_normal((n, d)).astype(np.float16)
    wq = rng.integers(-127, 128, size=(o, d), dtype=np.int8)
    s_w = ((1.0 + np.arange(o) / o) / (127.0 * np.sqrt(d))).astype(np.float32)
    b = rng.standard_normal(o).astype(np.float16)
    got = b200.run(DT_F16, (n, d), [Layer(H_FC, (n, o), w=wq, b=b, s_w=s_w)], x, run_mode=RM_GRAPH)
    want = x.astype(np.float32) @ (wq.astype(np.float32) * s_w.reshape(-1, 1)).T + b.astype(np.float32)
    f16_close(got, want)  # exact integer weights in fp16 + per-channel scale in the f32 epilogue: the fp16 path's 1e-3


from test_oracle import SE_CASES, se_case


@pytest.mark.gpu
@pytest.mark.parametrize("kind,op", SE_CASES, ids=["add", "mul", "sub"])
def test_binary_ops_between_an_activation_and_a_per_image_per_channel_activation(kind, op, b200, oracle, rng):
    """[N, C, H, W] (op) [N, C, 1, 1] with both operands produced by earlier layers (a squeeze-and-excitation scale): the second
    operand is indexed (image, channel) -- b200_binary_bcast_nc; int8 bit-exact in graph mode, fp16 within tolerance"""
    for shape in ((2, 24, 5, 7), (3, 64, 9, 9)):
        x, layers, want = se_case(kind, op, False, oracle, rng, shape=shape)
        got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.04, zp_in=3, run_mode=RM_GRAPH)
        assert np.array_equal(got, want), (shape, int(np.count_nonzero(got != want)))
    from shl import H_GAP
    shape = (2, 24, 5, 7)
    xh = rng.standard_normal(shape).astype(np.float16)
    layers = [Layer(H_RELU, shape), Layer(H_GAP, (2, 24, 1, 1), kernel=(5, 7)), Layer(kind, shape, in0=1, in1=2)]
    got = b200.run(DT_F16, shape, layers, xh, run_mode=RM_GRAPH)
    r = np.maximum(xh.astype(np.float32), 0)
    g = r.mean(axis=(2, 3), keepdims=True).astype(np.float16).astype(np.float32)
    f16_close(got, {0: r + g, 1: r - g, 2: r * g}[op])


from test_oracle import SPLIT_CASES, split_case


@pytest.mark.gpu
@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("shape,axis,at", SPLIT_CASES)
def test_split_int8_bit_exact(shape, axis, at, which, b200, oracle, rng):
    x, layers, want = split_case(shape, axis, at, which, oracle, rng)
    for mode in (RM_LAYER, RM_GRAPH):
        got = b200.run(DT_INT8, shape, layers, x, s_in=0.04, zp_in=3, run_mode=mode)
        assert np.array_equal(got, want), mode


@pytest.mark.gpu
def test_split_fp16_is_a_copy(b200, rng):
    from shl import H_SPLIT
    for shape, axis, at in SPLIT_CASES:
        x = rng.standard_normal(shape).astype(np.float16)
        for which in (0, 1):
            out_shape = list(shape)
            out_shape[axis] = at if which == 0 else shape[axis] - at
            layers = [Layer(H_RELU, shape), Layer(H_SPLIT, tuple(out_shape), axis=axis, p0=float(at), p1=float(which))]
            got = b200.run(DT_F16, shape, layers, x, run_mode=RM_GRAPH)
            sl = [slice(None)] * 4
            sl[axis] = slice(0, at) if which == 0 else slice(at, None)
            want = np.maximum(x, np.float16(0))[tuple(sl)]
            assert np.array_equal(got.view(np.uint16), np.ascontiguousarray(want).view(np.uint16)), (shape, axis, which)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
def test_shufflenet_style_unit_int8_bit_exact(mode, b200, oracle, rng):
    """split -> (identity branch | 1x1 conv + relu -> depthwise 3x3 -> 1x1 conv) -> concat, the ShuffleNetV2 basic
    unit without the shuffle: multi-output and multi-input nodes share tensors in the planned arena"""
    from shl import H_CONCAT, H_SPLIT
    n, c, h, w = 3, 48, 14, 14
    half = c // 2
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    w1, sw1, b1, so1 = synth_conv_i8(rng, half, half, 1, 1)
    wd, swd, bd, sod = synth_conv_i8(rng, half, half, 3, 3, depthwise=True)
    w2, sw2, b2, so2 = synth_conv_i8(rng, half, half, 1, 1)
    layers = [
        Layer(H_RELU, (n, c, h, w), s_out=0.02, zp_out=-128),                                               # t1
        Layer(H_SPLIT, (n, half, h, w), in0=1, s_out=0.02, zp_out=-128, axis=1, p0=float(half), p1=0.0),   # t2
        Layer(H_SPLIT, (n, half, h, w), in0=1, s_out=0.02, zp_out=-128, axis=1, p0=float(half), p1=1.0),   # t3
        Layer(H_CONV_RELU, (n, half, h, w), in0=3, w=w1, b=b1, s_w=sw1, s_out=so1, zp_out=-128),            # t4
        Layer(H_CONV, (n, half, h, w), in0=4, w=wd, b=bd, s_w=swd, s_out=sod, zp_out=5, pad=(1,) * 4, group=half),  # t5
        Layer(H_CONV_RELU, (n, half, h, w), in0=5, w=w2, b=b2, s_w=sw2, s_out=so2, zp_out=-128),            # t6
        Layer(H_CONCAT, (n, c, h, w), in0=2, in1=6, s_out=0.03, zp_out=-128, axis=1),                        # t7
    ]
    got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.04, zp_in=3, run_mode=mode)
    t1 = oracle.relu_i8(x, ACT_RELU, 0.04, 3, 0.02, -128)
    t2, t3 = np.ascontiguousarray(t1[:, :half]), np.ascontiguousarray(t1[:, half:])   # same qinfo: identity tables
    kw = dict(stride=(1, 1), dilation=(1, 1), group=1, s_b=None)
    t4 = oracle.conv2d_i8(t3, w1, b1, t3.shape, pad=(0,) * 4, s_in=0.02, zp_in=-128, s_w=sw1, s_out=so1, zp_out=-128,
                          act=ACT_RELU, **kw)
    t5 = oracle.conv2d_i8(t4, wd, bd, t4.shape, depthwise=True, pad=(1,) * 4, s_in=so1, zp_in=-128, s_w=swd, s_out=sod,
                          zp_out=5, **kw)
    t6 = oracle.conv2d_i8(t5, w2, b2, t5.shape, pad=(0,) * 4, s_in=sod, zp_in=5, s_w=sw2, s_out=so2, zp_out=-128,
                          act=ACT_RELU, **kw)
    want = oracle.concat_i8([t2, t6], [(0.02, -128), (so2, -128)], 1, 0.03, -128)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)}/{got.size} differ"


@pytest.mark.gpu
def test_swizzled_descriptor_start_may_be_shifted_by_any_number_of_rows():
    """hardware fact the next kernels build on (csrc/umma_probe.cu, tools/umma_probe.py): a SWIZZLE_128B K-major A
    descriptor whose start address is `shift` 128-byte rows into a TMA-written tile reads rows shift .. shift+127 for
    EVERY shift, not only multiples of the 8-row swizzle period -- the swizzle follows the absolute shared-memory
    address.  So the taps of a 3x3 convolution can be read from one flat halo tile by shifting the descriptor."""
    import ctypes as C
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = C.CDLL(os.path.join(root, "csi-nn2_b200", "lib", "libb200nn.so"))
    shim.b200_last_error.restype = C.c_char_p
    rng = np.random.default_rng(0)
    a = rng.integers(-128, 128, size=(144, 128), dtype=np.int8)
    b = rng.integers(-128, 128, size=(32, 128), dtype=np.int8)
    d_a, d_b, d_o = C.c_void_p(), C.c_void_p(), C.c_void_p()
    for d, n in ((d_a, a.nbytes), (d_b, b.nbytes), (d_o, 128 * 32 * 4)):
        assert shim.b200_malloc(C.byref(d), C.c_size_t(n)) == 0, shim.b200_last_error()
    assert shim.b200_memcpy_h2d(d_a, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), None) == 0
    assert shim.b200_memcpy_h2d(d_b, b.ctypes.data_as(C.c_void_p), C.c_size_t(b.nbytes), None) == 0
    for k0 in range(4):
        for shift in (0, 1, 2, 3, 5, 7, 8, 9, 13, 16):
            got = np.zeros((128, 32), np.int32)
            assert shim.b200_test_umma_shifted_start(d_a, d_b, shift, k0, d_o, None) == 0, shim.b200_last_error()
            assert shim.b200_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_o, C.c_size_t(got.nbytes), None) == 0
            assert shim.b200_stream_sync(None) == 0, shim.b200_last_error()
            ks = slice(32 * k0, 32 * k0 + 32)
            want = a[shift:shift + 128, ks].astype(np.int32) @ b[:, ks].astype(np.int32).T
            assert np.array_equal(got, want), (k0, shift)
    for d in (d_a, d_b, d_o):
        shim.b200_free(d)


@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
def test_flatten_of_a_feature_map_feeds_fullyconnected(mode, b200, oracle, rng):
    """N x C x H x W -> N x (C*H*W) (the flatten in front of a VGG / AlexNet classifier, source/reference/flatten.c):
    the pixel-major device tensor is permuted back into NCHW order, then the fullyconnected GEMM runs on the rows"""
    n, c, h, w, units = 3, 32, 5, 4, 40
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt4, s_w, b, s_out = synth_conv_i8(rng, c * h * w, units, 1, 1)
    layers = [Layer(H_RELU, (n, c, h, w), s_out=0.02, zp_out=-3),
              Layer(H_FLATTEN, (n, c * h * w), s_out=0.02, zp_out=-3),
              Layer(H_FC, (n, units), s_out=s_out, zp_out=2, w=wt4.reshape(units, c * h * w), b=b, s_w=s_w)]
    from types import SimpleNamespace
    want = nets.oracle_forward(SimpleNamespace(layers=layers, orc=oracle, dtype=DT_INT8, s_in=0.02, zp_in=-3), x)
    got = b200.run(DT_INT8, x.shape, layers, x, s_in=0.02, zp_in=-3, run_mode=mode)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
def test_general_reshape_keeps_the_row_major_bytes(mode, b200, rng):
    """csinn_reshape between shapes that are neither N x C x 1 x 1 <-> N x C nor the flatten (source/reference/reshape.c
    copies the NCHW bytes): on the device the tensor goes through NCHW order in a scratch buffer and back into the
    pixel-major layout of the new shape; the result read back must be the input's bytes under the new shape"""
    from shl import H_RESHAPE
    for shape, new in (((2, 24, 5, 6), (2, 8, 15, 6)), ((1, 6, 4, 10), (1, 40, 2, 3)), ((3, 16, 2, 2), (3, 4, 4, 4)),
                       ((2, 12, 3, 7), (2, 252))):
        x = rng.integers(-128, 128, size=shape, dtype=np.int8)
        layers = [Layer(H_RELU, shape, s_out=0.02, zp_out=-128), Layer(H_RESHAPE, new, s_out=0.02, zp_out=-128)]
        got = b200.run(DT_INT8, shape, layers, x, s_in=0.02, zp_in=-128, run_mode=mode)
        assert np.array_equal(got, np.maximum(x, -128).reshape(new)), (shape, new)
        xh = rng.standard_normal(shape).astype(np.float16)
        goth = b200.run(DT_F16, shape, [Layer(H_RELU, shape), Layer(H_RESHAPE, new)], xh, run_mode=mode)
        assert np.array_equal(goth.view(np.uint16), np.maximum(xh, np.float16(0)).reshape(new).view(np.uint16)), (shape, new)
