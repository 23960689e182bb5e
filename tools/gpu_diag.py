#!/usr/bin/env python
"""Developer diagnostic (not a test): run a ladder of b200 cases through the CSI-NN2 API on the
GPU and print, per case, how the result differs from the oracle -- without stopping at the first
failure, so that one gpurun call localises a kernel bug (which rows / columns / k-blocks).

    gpurun -- 'timeout 600 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1'
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from shl import *  # noqa: E402,F403

rng = np.random.default_rng(0)
b200 = Harness("b200")
orc = Oracle()
fails = 0


def report(name, got, want, tol=None):
    global fails
    if tol is None:
        d = got.astype(np.int64) - want.astype(np.int64)
        bad = np.count_nonzero(d)
    else:
        g, w = got.astype(np.float32), want.astype(np.float32)
        d = np.abs(g - w) / np.maximum(np.abs(w), 1.0)
        bad = np.count_nonzero(d > tol)
    status = "OK " if bad == 0 else "BAD"
    print(f"[{status}] {name}: mismatches {bad}/{got.size} max|d| {np.abs(d).max():.4g}", flush=True)
    if bad:
        fails += 1
        idx = np.argwhere(d != 0 if tol is None else d > tol)
        print("      first bad indices:", idx[:6].tolist())
        for ax in range(got.ndim):
            u = np.unique(idx[:, ax])
            print(f"      axis {ax}: {len(u)} distinct bad coords, e.g. {u[:12].tolist()}")
        flat = tuple(idx[0])
        print("      got", got[flat], "want", want[flat])


def conv_i8(name, n, c, h, w, o, k, stride=1, pad=0, group=1, dw=False, zp_in=0, kind=H_CONV, mode=RM_LAYER,
            post=None):
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k, group=group, depthwise=dw)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layers = [Layer(kind, (n, o, oh, ow), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w, stride=(stride, stride),
                    pad=(pad,) * 4, group=(c if dw else group))]
    act = {H_CONV: ACT_NONE, H_CONV_RELU: ACT_RELU, H_CONV_RELU6: ACT_RELU6}.get(kind, ACT_NONE)
    pst = None
    if post is not None:
        layers.append(Layer(H_RELU, (n, o, oh, ow), s_out=post[0], zp_out=post[1]))
        pst = (ACT_RELU, post[0], post[1])
    t0 = time.time()
    try:
        got = b200.run(DT_INT8, (n, c, h, w), layers, x, s_in=0.02, zp_in=zp_in, run_mode=mode)
    except Exception as e:  # noqa: BLE001
        global fails
        fails += 1
        print(f"[ERR] {name}: {e}", flush=True)
        return
    want = orc.conv2d_i8(x, wt, b, (n, o, oh, ow), depthwise=dw, stride=(stride, stride), pad=(pad,) * 4,
                         dilation=(1, 1), group=group, s_in=0.02, zp_in=zp_in, s_w=s_w, s_b=None, s_out=s_out,
                         zp_out=3, act=act, post=pst)
    report(f"{name} ({time.time() - t0:.2f}s)", got, want)


def conv_f16(name, n, c, h, w, o, k, stride=1, pad=0, dw=False, mode=RM_LAYER):
    x = rng.standard_normal((n, c, h, w)).astype(np.float16)
    cg = 1 if dw else c
    wt = (rng.standard_normal((o, cg, k, k)) / np.sqrt(cg * k * k)).astype(np.float16)
    b = rng.standard_normal(o).astype(np.float16)
    oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
    layers = [Layer(H_CONV, (n, o, oh, ow), w=wt, b=b, stride=(stride, stride), pad=(pad,) * 4, group=(c if dw else 1))]
    try:
        got = b200.run(DT_F16, (n, c, h, w), layers, x, run_mode=mode)
    except Exception as e:  # noqa: BLE001
        global fails
        fails += 1
        print(f"[ERR] {name}: {e}", flush=True)
        return
    want = orc.conv2d_f32(x.astype(np.float32), wt.astype(np.float32), b.astype(np.float32), (n, o, oh, ow),
                          depthwise=dw, stride=(stride, stride), pad=(pad,) * 4)
    report(name, got, want, tol=2e-3)


if __name__ == "__main__":
    print("harness loaded; running ladder", flush=True)
    # 1x1 conv == plain GEMM: one tile, then tails in M, N, K, then multi-tile / multi-k-block
    conv_i8("gemm 128x64x128", 1, 128, 8, 16, 64, 1)
    conv_i8("gemm 128x64x32 (K<128)", 1, 32, 8, 16, 64, 1)
    conv_i8("gemm 128x16x16 (min)", 1, 16, 8, 16, 16, 1)
    conv_i8("gemm M tail 100", 1, 64, 10, 10, 64, 1)
    conv_i8("gemm N tail 40", 1, 64, 8, 16, 40, 1)
    conv_i8("gemm K tail 72", 1, 72, 8, 16, 64, 1)
    conv_i8("gemm K 1024 (8 k-blocks)", 1, 1024, 7, 7, 128, 1)
    conv_i8("gemm N 1000 (4 n-tiles)", 2, 1024, 1, 1, 1000, 1)
    conv_i8("gemm many tiles", 2, 128, 56, 56, 128, 1)
    conv_i8("gemm c not x16 (C=20,O=24)", 1, 20, 9, 9, 24, 1)
    conv_i8("gemm zp_in=-7", 1, 64, 14, 14, 64, 1, zp_in=-7)
    conv_i8("gemm conv2d_relu", 1, 64, 14, 14, 64, 1, kind=H_CONV_RELU)
    conv_i8("gemm conv2d_relu6", 1, 64, 14, 14, 64, 1, kind=H_CONV_RELU6)
    # im2col path
    conv_i8("conv3x3 s1 p1", 1, 32, 14, 14, 48, 3, 1, 1, zp_in=-7)
    conv_i8("conv3x3 s2 p1 C=3", 2, 3, 32, 32, 32, 3, 2, 1, zp_in=5)
    conv_i8("conv7x7 s2 p3", 1, 3, 40, 40, 64, 7, 2, 3)
    conv_i8("group conv g=4", 1, 64, 12, 12, 64, 3, 1, 1, group=4)
    # depthwise
    conv_i8("dw3x3 s1", 2, 32, 20, 20, 32, 3, 1, 1, dw=True, zp_in=-7)
    conv_i8("dw3x3 s2", 1, 64, 21, 21, 64, 3, 2, 1, dw=True)
    conv_i8("dw3x3 C=24", 1, 24, 9, 11, 24, 3, 1, 1, dw=True)
    conv_i8("dw5x5", 1, 16, 12, 12, 16, 5, 1, 2, dw=True)
    # graph mode incl. relu fusion with its own qinfo
    conv_i8("graph conv3x3(nchw in)+relu", 2, 3, 32, 32, 32, 3, 2, 1, mode=RM_GRAPH, post=(0.05, -128))
    conv_i8("graph 1x1+relu", 1, 64, 14, 14, 64, 1, mode=RM_GRAPH, post=(0.04, -100))
    conv_i8("graph dw+relu", 1, 32, 14, 14, 32, 3, 1, 1, dw=True, mode=RM_GRAPH, post=(0.04, -100))
    # fp16
    conv_f16("f16 gemm 128x64x64", 1, 64, 8, 16, 64, 1)
    conv_f16("f16 gemm K=512 N=200", 1, 512, 9, 9, 200, 1)
    conv_f16("f16 conv3x3 s2 C=3", 1, 3, 32, 32, 32, 3, 2, 1)
    conv_f16("f16 dw3x3", 1, 32, 14, 14, 32, 3, 1, 1, dw=True)
    conv_f16("f16 graph conv3x3", 1, 3, 32, 32, 32, 3, 2, 1, mode=RM_GRAPH)
    print("DONE, failing cases:", fails, flush=True)
    sys.exit(1 if fails else 0)
