"""GPU tests of the implicit-GEMM convolution path (csrc/conv_igemm.cu) and of the hardware facts it rests on.

test_tma_im2col_semantics pins the conventions of TMA's im2col mode (cuTensorMapEncodeIm2col +
cp.async.bulk.tensor.4d...im2col) against a numpy gather: corner order, base-pixel coordinates, the walk over
W -> H -> N with the convolution stride, tap offsets, zero fill outside the image.
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shim():
    shim = C.CDLL(os.path.join(ROOT, "csi-nn2_b200", "lib", "libb200nn.so"))
    shim.b200_last_error.restype = C.c_char_p
    return shim


def im2col_rows(x, kh, kw, stride, pads, dil, ky, kx, m0, pixels, c0, chans):
    """[pixels][chans]: output pixels m0.. (n, oy, ox order) of tap (ky, kx), channels c0.., zeros outside the image"""
    n, h, w, c = x.shape
    pt, pl, pb, pr = pads
    oh = (h + pt + pb - dil * (kh - 1) - 1) // stride + 1
    ow = (w + pl + pr - dil * (kw - 1) - 1) // stride + 1
    out = np.zeros((pixels, chans), x.dtype)
    for i in range(pixels):
        m = m0 + i
        b, r = divmod(m, oh * ow)
        oy, ox = divmod(r, ow)
        iy, ix = oy * stride - pt + ky * dil, ox * stride - pl + kx * dil
        if b < n and 0 <= iy < h and 0 <= ix < w:
            cc = min(chans, c - c0)
            out[i, :cc] = x[b, iy, ix, c0:c0 + cc]
    return out, oh, ow


@pytest.mark.parametrize("case", [
    # n, h, w, c, cp, kh, kw, stride, (pt, pl, pb, pr), dil, chans, pixels
    (3, 9, 11, 64, 64, 3, 3, 1, (1, 1, 1, 1), 1, 64, 128),
    (2, 14, 14, 128, 128, 3, 3, 2, (1, 1, 1, 1), 1, 128, 128),
    (2, 13, 10, 64, 80, 3, 3, 1, (0, 1, 1, 2), 1, 64, 96),      # every pad different: pins the corner order
    (2, 12, 12, 32, 32, 3, 3, 1, (2, 2, 2, 2), 2, 32, 128),      # dilation 2
    (2, 15, 9, 64, 64, 5, 3, 2, (2, 1, 2, 1), 1, 64, 64),        # 5 x 3 kernel, stride 2
    (4, 7, 7, 256, 256, 3, 3, 1, (1, 1, 1, 1), 1, 128, 128),     # a tile spans several images; second channel slab
], ids=lambda c: "n%d_%dx%d_c%d_cp%d_k%dx%d_s%d_p%s_d%d_ch%d_px%d" % (c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7],
                                                                    "".join(map(str, c[8])), c[9], c[10], c[11]))
def test_tma_im2col_semantics(case):
    n, h, w, c, cp, kh, kw, stride, pads, dil, chans, pixels = case
    pt, pl, pb, pr = pads
    shim = _shim()
    rng = np.random.default_rng(7)
    x = rng.integers(1, 128, size=(n, h, w, cp), dtype=np.int8)  # non-zero: a zero must come from the fill
    d_x, d_o = C.c_void_p(), C.c_void_p()
    assert shim.b200_malloc(C.byref(d_x), C.c_size_t(x.nbytes)) == 0
    assert shim.b200_malloc(C.byref(d_o), C.c_size_t(pixels * chans)) == 0
    assert shim.b200_memcpy_h2d(d_x, x.ctypes.data_as(C.c_void_p), C.c_size_t(x.nbytes), None) == 0
    lower_w, lower_h = -pl, -pt
    upper_w, upper_h = pr - (kw - 1) * dil, pb - (kh - 1) * dil
    _, oh, ow = im2col_rows(x[..., :c], kh, kw, stride, pads, dil, 0, 0, 0, 1, 0, chans)
    total = n * oh * ow
    for (ky, kx, m0, c0) in [(0, 0, 0, 0), (1, 1, 0, 0), (kh - 1, kw - 1, 0, 0), (0, kw - 1, ow - 3, 0),
                             (kh - 1, 0, max(0, oh * ow - 5), 0), (1, 0, max(0, total - pixels), 0)] + \
            ([(1, 1, 5, 128)] if c > chans else []):
        want, _, _ = im2col_rows(x[..., :c], kh, kw, stride, pads, dil, ky, kx, m0, pixels, c0, chans)
        b0, r = divmod(m0, oh * ow)
        oy0, ox0 = divmod(r, ow)
        got = np.zeros((pixels, chans), np.int8)
        rc = shim.b200_test_tma_im2col(d_x, n, h, w, c, cp, lower_w, lower_h, upper_w, upper_h, stride, stride, chans, pixels,
                                       c0, ox0 * stride + lower_w, oy0 * stride + lower_h, b0, kx * dil, ky * dil, d_o, None)
        assert rc == 0, shim.b200_last_error()
        assert shim.b200_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_o, C.c_size_t(got.nbytes), None) == 0
        assert shim.b200_stream_sync(None) == 0, shim.b200_last_error()
        bad = np.argwhere(got != want)
        assert bad.size == 0, (ky, kx, m0, c0, len(bad), bad[:6].tolist(), got[bad[0][0], :8], want[bad[0][0], :8])
    shim.b200_free(d_x), shim.b200_free(d_o)


# ---- the implicit-GEMM convolution through the CSI-NN2 API ------------------------------------------------------
from shl import (ACT_RELU, DT_INT8, H_CONV, H_CONV_RELU, H_RELU, RM_GRAPH, RM_LAYER, Layer, conv_out_hw,  # noqa: E402
                 synth_conv_i8)

IGEMM_CASES = [
    # n, c, h, w, o, kh, kw, stride, pad, dil, zp_in
    (2, 64, 14, 14, 64, 3, 3, 1, 1, 1, -128),    # ResNet layer1 3x3: 64-byte K blocks (SWIZZLE_64B), 9 border classes
    (1, 64, 56, 56, 64, 3, 3, 1, 1, 1, -7),      # several tiles per CTA: the re-seeded classes of later tiles
    (2, 128, 28, 28, 128, 3, 3, 2, 1, 1, -128),  # stride 2, 128-byte K blocks
    (3, 128, 9, 11, 96, 3, 3, 1, 1, 1, 5),       # ragged M and N, tiles that span image borders
    (1, 256, 14, 14, 256, 3, 3, 1, 1, 1, -128),  # weights not resident (K * n-tile > 160 KB): streamed B stages
    (2, 64, 12, 12, 48, 3, 3, 1, 0, 1, 9),       # no padding: one class, plain ibias seeds
    (2, 64, 12, 12, 80, 3, 3, 1, 1, 1, 0),       # zp_in = 0: zero fill is already right
    (1, 64, 19, 17, 64, 3, 3, 1, 2, 2, -5),      # dilation 2
    (1, 128, 15, 13, 32, 5, 3, 2, 2, 2, 3),      # 5 x 3 kernel, stride 2, more border classes
    (5, 512, 7, 7, 512, 3, 3, 1, 1, 1, -128),    # ResNet layer4 3x3: 36 K blocks, 4 n-tiles
    (2, 256, 28, 28, 512, 1, 1, 2, 0, 1, -128),  # ResNet downsample shortcut: strided 1x1 (one tap, no im2col buffer either)
    (3, 64, 13, 13, 64, 1, 1, 2, 0, 1, 0),       # strided 1x1, odd size, 64-byte K blocks
]


@pytest.mark.parametrize("case", IGEMM_CASES, ids=lambda c: "n%d_c%d_%dx%d_o%d_k%dx%d_s%d_p%d_d%d_zp%d" % c)
@pytest.mark.parametrize("mode", [RM_LAYER, RM_GRAPH], ids=["layer", "graph"])
def test_conv_igemm_bit_exact(case, mode, b200, oracle, rng):
    n, c, h, w, o, kh, kw, stride, pad, dil, zp_in = case
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt = rng.integers(-127, 128, size=(o, c, kh, kw), dtype=np.int8)
    _, s_w, b, s_out = synth_conv_i8(rng, c, o, kh, kw)
    oh, ow = conv_out_hw(h, w, kh, kw, (stride, stride), (pad,) * 4, (dil, dil))
    pre = Layer(H_RELU, (n, c, h, w), s_out=0.02, zp_out=zp_in)  # a producer, so that the conv reads a pixel-major tensor
    conv = Layer(H_CONV, (n, o, oh, ow), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w, stride=(stride, stride), pad=(pad,) * 4,
                 dilation=(dil, dil))
    post = Layer(H_RELU, (n, o, oh, ow), s_out=s_out / 2, zp_out=-128)
    kw_ = dict(stride=(stride, stride), pad=(pad,) * 4, dilation=(dil, dil), group=1, s_in=0.02, zp_in=zp_in, s_w=s_w,
               s_b=None, s_out=s_out, zp_out=3)
    if mode == RM_LAYER:
        got = b200.run(DT_INT8, x.shape, [conv], x, s_in=0.02, zp_in=zp_in)
        assert np.array_equal(got, oracle.conv2d_i8(x, wt, b, (n, o, oh, ow), **kw_))
        return
    xr = oracle.relu_i8(x, ACT_RELU, 0.02, zp_in, 0.02, zp_in)
    with b200.create(DT_INT8, x.shape, [pre, conv, post], s_in=0.02, zp_in=zp_in, run_mode=RM_GRAPH) as net:
        assert "b200_conv_igemm_tcgen05" in net.describe(), net.describe()
        got = net(x)
    want = oracle.conv2d_i8(xr, wt, b, (n, o, oh, ow), post=(ACT_RELU, s_out / 2, -128), **kw_)
    bad = np.argwhere(got != want)
    assert bad.size == 0, (len(bad), got.size, bad[:8].tolist())
    os.environ["SHL_B200_NO_IGEMM"] = "1"
    try:
        with b200.create(DT_INT8, x.shape, [pre, conv, post], s_in=0.02, zp_in=zp_in, run_mode=RM_GRAPH) as net:
            assert "b200_conv_igemm_tcgen05" not in net.describe()
            assert np.array_equal(net(x), got), "implicit GEMM != im2col + GEMM"
    finally:
        os.environ.pop("SHL_B200_NO_IGEMM", None)
