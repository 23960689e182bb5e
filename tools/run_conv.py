#!/usr/bin/env python
"""One csinn_conv2d (behind a relu node, its own relu fused) as its own graph-mode session: per-step device time, TOPS and
GB/s -- the shapes behind the bench line's conv2d_tops, and a small target for `ncu --set full`.

    python tools/run_conv.py N C HW O K [stride] [reps]          # SHL_B200_NO_IGEMM=1: explicit im2col + GEMM
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from shl import DT_INT8, H_CONV, H_RELU, RM_GRAPH, Harness, Layer, conv_out_hw, synth_conv_i8  # noqa: E402


def main():
    n, c, hw, o, k = (int(v) for v in sys.argv[1:6])
    stride = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 10
    rng = np.random.default_rng(3)
    oh, ow = conv_out_hw(hw, hw, k, k, (stride, stride), (k // 2,) * 4)
    wt = rng.integers(-127, 128, size=(o, c, k, k), dtype=np.int8)
    _, s_w, b, s_out = synth_conv_i8(rng, c, o, k, k)
    layers = [Layer(H_RELU, (n, c, hw, hw), s_out=0.02, zp_out=-128),
              Layer(H_CONV, (n, o, oh, ow), s_out=s_out, zp_out=0, w=wt, b=b, s_w=s_w, stride=(stride, stride), pad=(k // 2,) * 4),
              Layer(H_RELU, (n, o, oh, ow), s_out=s_out / 2, zp_out=-128)]
    x = rng.integers(-128, 128, size=(n, c, hw, hw), dtype=np.int8)
    b200 = Harness("b200")
    shl = C.CDLL(os.path.join(ROOT, "csi-nn2_b200", "lib", "libshl_b200.so"))
    shl.shl_b200_session_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), C.c_int]
    with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=-128, run_mode=RM_GRAPH) as net:
        net(x)
        cap = 8
        ms, by, op = (C.c_double * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
        kk = shl.shl_b200_session_profile(net.session, 3, reps, ms, by, op, cap)
        print(net.describe().strip())
        for i in range(kk):
            print(f"step {i}: {ms[i] * 1e3:8.1f} us  {by[i] / ms[i] / 1e6:7.0f} GB/s  {op[i] / ms[i] / 1e9:7.1f} TOPS")


if __name__ == "__main__":
    main()
