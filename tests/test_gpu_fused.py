"""GPU parity of the fused MobileNet block (csrc/dwpw_fused.cu): a depthwise 3x3 whose
only reader is a 1x1 convolution runs as ONE kernel in graph mode; the result must equal the oracle
chain bit for bit and the two-kernel path (SHL_B200_NO_DWPW=1) byte for byte.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest

import nets
from shl import (DT_INT8, H_CONV, H_CONV_RELU, H_CONV_RELU6, H_RELU, H_RELU6, RM_GRAPH, Layer, Oracle, conv_out_hw,
                 synth_conv_i8)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _always_fuse():
    """the planner normally times the fused kernel against the two-kernel path and keeps the faster
    (b200_dwpw_prefers_fusion); these tests are about the fused kernel, so they force it"""
    os.environ["SHL_B200_DWPW"] = "1"
    yield
    os.environ.pop("SHL_B200_DWPW", None)

# n, c, h, w, o, stride, pad, zp_in
PAIRS = [
    (2, 32, 112, 112, 64, 1, 1, 0),        # MobileNetV1 block 1: 32-byte K, two M128 blocks per tile
    (2, 64, 112, 112, 128, 2, 1, -128),    # block 2: stride 2
    (2, 128, 56, 56, 128, 1, 1, -128),     # block 3
    (2, 128, 56, 56, 256, 2, 1, -128),     # block 4: 256 output columns (two staging halves)
    (2, 256, 28, 28, 256, 1, 1, -128),     # block 5: two K blocks, several halo chunks per tile
    (3, 16, 9, 11, 24, 1, 1, -7),          # ragged everything
    (1, 48, 15, 15, 40, 2, 1, 4),          # channels not a multiple of 32 / 16-column tail
    (2, 96, 30, 33, 200, 1, 1, -5),        # ragged 256-column tile
    (5, 256, 7, 7, 256, 1, 1, 3),          # maps smaller than a tile
    (1, 128, 9, 10, 64, 2, 0, 2),          # unpadded, stride 2
    (1, 20, 6, 5, 16, 1, 0, 1),            # unpadded, tiny
    (150, 128, 14, 14, 128, 1, 1, -6),     # more tiles than SMs: every pipeline wraps
    (1, 160, 57, 9, 144, 2, 1, 7),         # 144 outputs: one full and one 16-column staging half
]


def _pair_layers(rng, n, c, h, w, o, stride, pad, mode):
    """[depthwise 3x3 (+ act), pointwise 1x1 (+ act)] with synthetic per-channel parameters; `mode`:
    'relu_own_qinfo' = standalone relu nodes with their own qinfo (the example graph: table epilogues),
    'relu_keeps_qinfo' = relu nodes that keep the producer's qinfo (clamp epilogues), 'conv_relu6' =
    csinn_conv2d_relu6 front ends, 'none' = no activation"""
    oh, ow = conv_out_hw(h, w, 3, 3, (stride, stride), (pad,) * 4)
    wd, s_wd, bd, s_d = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
    kind = H_CONV_RELU6 if mode == "conv_relu6" else H_CONV
    layers = [Layer(kind, (n, c, oh, ow), s_out=s_d, zp_out=3, w=wd, b=bd, s_w=s_wd, stride=(stride, stride),
                    pad=(pad,) * 4, group=c)]
    s_mid, zp_mid = s_d, 3
    if mode == "relu_own_qinfo":
        s_mid, zp_mid = s_d / 2, -128
        layers.append(Layer(H_RELU, (n, c, oh, ow), s_out=s_mid, zp_out=zp_mid))
    elif mode == "relu_keeps_qinfo":
        layers.append(Layer(H_RELU, (n, c, oh, ow), s_out=s_mid, zp_out=zp_mid))
    wp, s_wp, bp, s_p = synth_conv_i8(rng, c, o, 1, 1, s_in=s_mid)
    layers.append(Layer(kind, (n, o, oh, ow), s_out=s_p, zp_out=-5, w=wp, b=bp, s_w=s_wp))
    if mode == "relu_own_qinfo":
        layers.append(Layer(H_RELU6, (n, o, oh, ow), s_out=s_p / 2, zp_out=-128))
    elif mode == "relu_keeps_qinfo":
        layers.append(Layer(H_RELU, (n, o, oh, ow), s_out=s_p, zp_out=-5))
    return layers


def _oracle_chain(layers, x, s_in, zp_in):
    nb = SimpleNamespace(layers=layers, orc=Oracle(), dtype=DT_INT8, s_in=s_in, zp_in=zp_in)
    return nets.oracle_forward(nb, x)


@pytest.mark.parametrize("case", PAIRS, ids=lambda c: "n%d_c%d_%dx%d_o%d_s%d_p%d_zp%d" % c)
def test_fused_block_bit_exact(case, b200, rng):
    n, c, h, w, o, stride, pad, zp_in = case
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    for mode in ("relu_own_qinfo", "relu_keeps_qinfo") if n * h * w > 2000 else ("relu_own_qinfo", "relu_keeps_qinfo",
                                                                                  "conv_relu6", "none"):
        layers = _pair_layers(rng, n, c, h, w, o, stride, pad, mode)
        with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=zp_in, run_mode=RM_GRAPH) as net:
            assert "b200_dwpw_fused_tcgen05" in net.describe(), net.describe()
            got = net(x)
            again = net(x)
        want = _oracle_chain(layers, x, 0.02, zp_in)
        bad = np.argwhere(got != want)
        assert bad.size == 0, (mode, len(bad), got.size, bad[:8].tolist())
        assert np.array_equal(got, again), "replay differs"


def test_fused_block_equals_the_two_kernel_path(b200, rng):
    """the same graph with and without the fusion (SHL_B200_NO_DWPW=1): identical bytes, one step fewer"""
    n, c, h, w, o = 3, 64, 40, 44, 96
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    layers = _pair_layers(rng, n, c, h, w, o, 1, 1, "relu_own_qinfo")
    with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=-9, run_mode=RM_GRAPH) as net:
        d_fused = net.describe()
        fused = net(x)
    os.environ["SHL_B200_DWPW"] = "0"
    try:
        with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=-9, run_mode=RM_GRAPH) as net:
            d_plain = net.describe()
            plain = net(x)
    finally:
        os.environ["SHL_B200_DWPW"] = "1"
    assert "b200_dwpw_fused_tcgen05" in d_fused and "b200_dwpw_fused_tcgen05" not in d_plain
    assert d_fused.startswith("steps=1 ") and d_plain.startswith("steps=2 "), (d_fused, d_plain)
    assert np.array_equal(fused, plain)


def test_depthwise_with_a_second_reader_is_not_fused(b200, rng):
    """the depthwise result has a second reader (an add): the intermediate tensor must exist in HBM, so
    the planner keeps two kernels"""
    from shl import H_ADD
    n, c, h, w = 1, 32, 12, 12
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wd, s_wd, bd, s_d = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
    wp, s_wp, bp, s_p = synth_conv_i8(rng, c, c, 1, 1, s_in=s_d)
    layers = [Layer(H_CONV, (n, c, h, w), s_out=s_d, zp_out=0, w=wd, b=bd, s_w=s_wd, pad=(1,) * 4, group=c),
              Layer(H_CONV, (n, c, h, w), s_out=s_p, zp_out=0, w=wp, b=bp, s_w=s_wp),
              Layer(H_ADD, (n, c, h, w), in0=1, in1=2, s_out=s_p * 2, zp_out=0)]
    with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=0, run_mode=RM_GRAPH) as net:
        assert "b200_dwpw_fused_tcgen05" not in net.describe()
        got = net(x)
    assert np.array_equal(got, _oracle_chain(layers, x, 0.02, 0))
