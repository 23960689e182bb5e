#!/usr/bin/env python
"""One MobileNet block (depthwise 3x3 + act -> pointwise 1x1 + act) as its own graph-mode session: timing of the
fused kernel against the two-kernel path, and a small target for `ncu --set full`.

    python tools/run_pair.py N C H W O STRIDE [reps]        # SHL_B200_NO_DWPW=1 for the unfused pair
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from shl import DT_INT8, H_CONV, H_RELU, RM_GRAPH, Harness, Layer, conv_out_hw, synth_conv_i8  # noqa: E402


def main():
    n, c, h, w, o, stride = (int(v) for v in sys.argv[1:7])
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 20
    rng = np.random.default_rng(0)
    oh, ow = conv_out_hw(h, w, 3, 3, (stride, stride), (1,) * 4)
    wd, s_wd, bd, s_d = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
    wp, s_wp, bp, s_p = synth_conv_i8(rng, c, o, 1, 1, s_in=s_d / 2)
    layers = [Layer(H_CONV, (n, c, oh, ow), s_out=s_d, zp_out=0, w=wd, b=bd, s_w=s_wd, stride=(stride, stride), pad=(1,) * 4,
                    group=c),
              Layer(H_RELU, (n, c, oh, ow), s_out=s_d / 2, zp_out=-128),
              Layer(H_CONV, (n, o, oh, ow), s_out=s_p, zp_out=0, w=wp, b=bp, s_w=s_wp),
              Layer(H_RELU, (n, o, oh, ow), s_out=s_p / 2, zp_out=-128)]
    b200 = Harness("b200")
    shl = C.CDLL(os.path.join(ROOT, "csi-nn2_b200", "lib", "libshl_b200.so"))
    shl.shl_b200_session_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), C.c_int]
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=-128, run_mode=RM_GRAPH) as net:
        net(x)
        cap = 8
        ms, by, op = (C.c_double * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
        k = shl.shl_b200_session_profile(net.session, 3, reps, ms, by, op, cap)
        print(net.describe().strip())
        tot = sum(ms[i] for i in range(k))
        for i in range(k):
            print(f"step {i}: {ms[i] * 1e3:8.1f} us  {by[i] / ms[i] / 1e6:7.0f} GB/s  {op[i] / ms[i] / 1e9:7.1f} TOPS")
        print(f"total {tot * 1e3:.1f} us")


if __name__ == "__main__":
    main()
