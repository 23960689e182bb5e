#!/usr/bin/env python
"""Generate tests/golden/unit_kat.npz from the reference's own known-answer vectors.

Run in the build container (needs /root/reference; the tests only read the committed .npz):

    python tests/golden/make_golden.py

Source: /root/reference/tests/unit_test/valid_data/*.dat -- C byte arrays that the reference's
kernel-level tests feed to its back-end kernels (tests/unit_test/conv2d_im2col_gemm.c:118-139,
conv2d_1x1s1_gemm.c, dwconv2d.c:87-111, fullyconnected.c:90-111, maxpool.c, avgpool.c,
activation.c).  Shapes and parameters below are the ones those .c files pass.  The int8 arrays of
conv / dwconv / fc are empty `{}` in the reference, so int8 parity of the
contraction ops is pinned by executing the reference itself (tests/test_oracle.py), and these
vectors pin fp32 / fp16 (and int8 for maxpool / relu).  The `_ker1` / `_weight_ref` arrays are
RVV-reordered weights (backend specific) and are not taken.
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("REF", "/root/reference")
DATA = os.path.join(REF, "tests", "unit_test", "valid_data")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unit_kat.npz")

ARRAY = re.compile(r"unsigned char (\w+)\[\]\s*=\s*\{([^}]*)\};", re.S)


def load(fname):
    text = open(os.path.join(DATA, fname)).read()
    out = {}
    for name, body in ARRAY.findall(text):
        vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", body)]
        out[name] = np.array(vals, dtype=np.uint8)
    return out


def view(raw, dtype, shape):
    a = raw.view(dtype)
    assert a.size == int(np.prod(shape)), (a.size, shape)
    return a.reshape(shape).copy()


def main():
    if not os.path.isdir(DATA):
        sys.exit(f"{DATA} not found: run this where the reference tree is mounted")
    g = {}
    conv = load("conv2d.dat")
    dw = load("dwconv2d.dat")
    fc = load("fullyconnected.dat")
    mp = load("maxpool.dat")
    ap = load("avgpool.dat")
    act = load("activation.dat")
    for tag, dt in (("fp32", np.float32), ("fp16", np.float16)):
        # conv2d1x1s1: in [1,16,4,5] k [19,16,1,1] out [1,19,4,5] pad 0 (conv2d.dat:5-10)
        g[f"conv1x1_{tag}_in"] = view(conv[f"conv2d1x1s1_{tag}_in"], dt, (1, 16, 4, 5))
        g[f"conv1x1_{tag}_ker"] = view(conv[f"conv2d1x1s1_{tag}_ker"], dt, (19, 16, 1, 1))
        g[f"conv1x1_{tag}_bias"] = view(conv[f"conv2d1x1s1_{tag}_bias"], dt, (19,))
        g[f"conv1x1_{tag}_out"] = view(conv[f"conv2d1x1s1_{tag}_out"], dt, (1, 19, 4, 5))
        # conv2d_im2col: in [1,3,4,5] k [19,3,3,3] pad 1 stride 1 (conv2d.dat:523-527)
        g[f"conv3x3_{tag}_in"] = view(conv[f"conv2d_im2col_{tag}_in"], dt, (1, 3, 4, 5))
        g[f"conv3x3_{tag}_ker"] = view(conv[f"conv2d_im2col_{tag}_ker"], dt, (19, 3, 3, 3))
        g[f"conv3x3_{tag}_bias"] = view(conv[f"conv2d_im2col_{tag}_bias"], dt, (19,))
        g[f"conv3x3_{tag}_out"] = view(conv[f"conv2d_im2col_{tag}_out"], dt, (1, 19, 4, 5))
        # depthwise 3x3 s1 p1: [2,4,10] -> [2,4,10]; s2 p1: [2,6,18] -> [2,3,9] (dwconv2d.dat:4-6,106-108)
        g[f"dw3x3s1_{tag}_in"] = view(dw[f"dwconv3x3s1_{tag}_in"], dt, (1, 2, 4, 10))
        g[f"dw3x3s1_{tag}_ker"] = view(dw[f"dwconv3x3s1_{tag}_ker"], dt, (2, 1, 3, 3))
        g[f"dw3x3s1_{tag}_bias"] = view(dw[f"dwconv3x3s1_{tag}_bias"], dt, (2,))
        g[f"dw3x3s1_{tag}_out"] = view(dw[f"dwconv3x3s1_{tag}_out"], dt, (1, 2, 4, 10))
        g[f"dw3x3s2_{tag}_in"] = view(dw[f"dwconv3x3s2_{tag}_in"], dt, (1, 2, 6, 18))
        g[f"dw3x3s2_{tag}_ker"] = view(dw[f"dwconv3x3s2_{tag}_ker"], dt, (2, 1, 3, 3))
        g[f"dw3x3s2_{tag}_bias"] = view(dw[f"dwconv3x3s2_{tag}_bias"], dt, (2,))
        g[f"dw3x3s2_{tag}_out"] = view(dw[f"dwconv3x3s2_{tag}_out"], dt, (1, 2, 3, 9))
        # fullyconnected: in_node 17, out_node 31 (fullyconnected.dat:4)
        g[f"fc_{tag}_in"] = view(fc[f"fc_{tag}_in"], dt, (1, 17))
        g[f"fc_{tag}_weight"] = view(fc[f"fc_{tag}_weight"], dt, (31, 17))
        g[f"fc_{tag}_bias"] = view(fc[f"fc_{tag}_bias"], dt, (31,))
        g[f"fc_{tag}_out"] = view(fc[f"fc_{tag}_out"], dt, (1, 31))
        # avgpool (avgpool.c: [c,h,w] in -> out, kernel, stride, pad)
        g[f"avgpool2x2s2_{tag}_in"] = view(ap[f"avgpool2x2s2_{tag}_in"], dt, (1, 2, 6, 18))
        g[f"avgpool2x2s2_{tag}_out"] = view(ap[f"avgpool2x2s2_{tag}_out"], dt, (1, 2, 3, 9))
        g[f"avgpool3x3s2_{tag}_in"] = view(ap[f"avgpool3x3s2_{tag}_in"], dt, (1, 2, 7, 19))
        g[f"avgpool3x3s2_{tag}_out"] = view(ap[f"avgpool3x3s2_{tag}_out"], dt, (1, 2, 3, 9))
        g[f"global_avgpool_{tag}_in"] = view(ap[f"global_avgpool_{tag}_in"], dt, (1, 3, 7, 7))
        g[f"global_avgpool_{tag}_out"] = view(ap[f"global_avgpool_{tag}_out"], dt, (1, 3, 1, 1))
        # maxpool (maxpool.c:68-100)
        g[f"maxpool2x2s2_{tag}_in"] = view(mp[f"maxpool2x2s2_{tag}_in"], dt, (1, 2, 6, 18))
        g[f"maxpool2x2s2_{tag}_out"] = view(mp[f"maxpool2x2s2_{tag}_out"], dt, (1, 2, 3, 9))
        g[f"maxpool3x3s2_p1_{tag}_in"] = view(mp[f"maxpool3x3s2_p1_{tag}_in"], dt, (1, 2, 6, 18))
        g[f"maxpool3x3s2_p1_{tag}_out"] = view(mp[f"maxpool3x3s2_p1_{tag}_out"], dt, (1, 2, 3, 9))
        # relu
        n = act[f"relu_{tag}_in"].view(dt).size
        g[f"relu_{tag}_in"] = view(act[f"relu_{tag}_in"], dt, (n,))
        g[f"relu_{tag}_out"] = view(act[f"relu_{tag}_out"], dt, (n,))
    # the int8 goldens the reference does hold
    g["maxpool2x2s2_int8_in"] = view(mp["maxpool2x2s2_int8_in"], np.int8, (1, 2, 6, 18))
    g["maxpool2x2s2_int8_out"] = view(mp["maxpool2x2s2_int8_out"], np.int8, (1, 2, 3, 9))
    g["maxpool3x3s2_p1_int8_in"] = view(mp["maxpool3x3s2_p1_int8_in"], np.int8, (1, 2, 6, 18))
    g["maxpool3x3s2_p1_int8_out"] = view(mp["maxpool3x3s2_p1_int8_out"], np.int8, (1, 2, 3, 9))
    g["maxpool3x3s1_p1_int8_in"] = view(mp["maxpool3x3s1_p1_int8_in"], np.int8, (1, 2, 3, 10))
    g["maxpool3x3s1_p1_int8_out"] = view(mp["maxpool3x3s1_p1_int8_out"], np.int8, (1, 2, 3, 10))
    n = act["relu_int8_in"].size
    g["relu_int8_in"] = view(act["relu_int8_in"], np.int8, (n,))
    g["relu_int8_out"] = view(act["relu_int8_out"], np.int8, (n,))
    for k in ("dwconv3x3s1_int8_in", "dwconv3x3s2_int8_in"):
        assert dw[k].size == 0, "reference now ships int8 depthwise goldens: add them here"
    assert fc["fc_int8_in"].size == 0
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT}: {len(g)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
