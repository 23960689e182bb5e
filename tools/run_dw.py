#!/usr/bin/env python
"""One int8 depthwise 3x3 (+ a fused relu table) as its own graph-mode session: device time and GB/s of the step, checked
against the oracle on one image -- a small target for `ncu --set full` and for A/B runs of the depthwise kernels
(SHL_B200_DW_IMMA=0: the dp4a kernel).

    python tools/run_dw.py N C HW [stride] [reps]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from shl import DT_INT8, H_CONV, H_RELU, RM_GRAPH, Harness, Layer, Oracle, synth_conv_i8  # noqa: E402


def main():
    n, c, hw = (int(v) for v in sys.argv[1:4])
    stride = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    rng = np.random.default_rng(0)
    x = rng.integers(-128, 128, size=(n, c, hw, hw), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
    oh = (hw + 2 - 3) // stride + 1
    layers = [Layer(H_RELU, (n, c, hw, hw), s_out=0.02, zp_out=-128),
              Layer(H_CONV, (n, c, oh, oh), s_out=s_out, zp_out=0, w=wt, b=b, s_w=s_w, stride=(stride, stride), pad=(1,) * 4, group=c),
              Layer(H_RELU, (n, c, oh, oh), s_out=s_out / 2, zp_out=-128)]
    b200, orc = Harness("b200"), Oracle()
    shl = C.CDLL(os.path.join(ROOT, "csi-nn2_b200", "lib", "libshl_b200.so"))
    shl.shl_b200_session_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), C.c_int]
    with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=-128, run_mode=RM_GRAPH) as net:
        got = net(x)
        x1 = orc.relu_i8(x[:1], 1, 0.02, -128, 0.02, -128)
        want = orc.conv2d_i8(x1, wt, b, (1, c, oh, oh), depthwise=True, stride=(stride, stride), pad=(1,) * 4, dilation=(1, 1),
                             group=1, s_in=0.02, zp_in=-128, s_w=s_w, s_b=None, s_out=s_out, zp_out=0, post=(1, s_out / 2, -128))
        assert np.array_equal(got[:1], want), "depthwise result differs from the oracle"
        ms, by, op = (C.c_double * 8)(), (C.c_double * 8)(), (C.c_double * 8)()
        k = shl.shl_b200_session_profile(net.session, 3, reps, ms, by, op, 8)
        for i in range(k):
            print(f"N={n} C={c} {hw}x{hw} s{stride} step {i}: {ms[i] * 1e3:8.1f} us  {by[i] / ms[i] / 1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
