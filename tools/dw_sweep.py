#!/usr/bin/env python
"""BASELINE.json configs[3]: depthwise_conv2d 3x3 int8 sweep, C in {32..1024}, H = W = 56, batch 128,
one B200 -- achieved HBM GB/s against the measured roofline, per channel count.

    gpurun -- 'python tools/dw_sweep.py > gpurun_out/dw_sweep.json'

Each point is a one-layer graph-mode session through the CSI-NN2 API (csinn_conv2d with
group == C -> the depthwise callback), checked bit-exactly against the oracle on one image, then
timed on the device with CUDA events (shl_b200_session_profile, 3 warm-up + 10 timed replays,
inputs resident in HBM).  Algorithmic bytes per SURVEY.md 8d: N*C*56*56*2 + 9C + 4C.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from shl import DT_INT8, H_CONV, H_RELU, RM_GRAPH, Harness, Layer, Oracle, synth_conv_i8  # noqa: E402


def main():
    batch = int(os.environ.get("DW_SWEEP_BATCH", "128"))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    shl = C.CDLL(os.path.join(ROOT, "csi-nn2_b200", "lib", "libshl_b200.so"))
    shl.shl_b200_session_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    b200, orc = Harness("b200"), Oracle()
    rng = np.random.default_rng(0)
    points = []
    for stride in (1, 2):
        for c in (32, 64, 128, 256, 512, 1024):
            x = rng.integers(-128, 128, size=(batch, c, 56, 56), dtype=np.int8)
            wt, s_w, b, s_out = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
            oh = (56 + 2 - 3) // stride + 1
            layers = [Layer(H_CONV, (batch, c, oh, oh), s_out=s_out, zp_out=0, w=wt, b=b, s_w=s_w, stride=(stride, stride),
                            pad=(1,) * 4, group=c), Layer(H_RELU, (batch, c, oh, oh), s_out=s_out / 2, zp_out=-128)]
            with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=-128, run_mode=RM_GRAPH) as net:
                got = net(x)
                want = orc.conv2d_i8(x[:1], wt, b, (1, c, oh, oh), depthwise=True, stride=(stride, stride), pad=(1,) * 4,
                                     dilation=(1, 1), group=1, s_in=0.02, zp_in=-128, s_w=s_w, s_b=None, s_out=s_out,
                                     zp_out=0, post=(1, s_out / 2, -128))
                assert np.array_equal(got[:1], want), (c, stride)
                ms, by, op = (C.c_double * 8)(), (C.c_double * 8)(), (C.c_double * 8)()
                n = shl.shl_b200_session_profile(net.session, 3, 10, ms, by, op, 8)
                assert n >= 1
                # step 0 is the NCHW -> pixel-major conversion of the graph input when present; the
                # depthwise step is the one with ops
                i = max(range(n), key=lambda k: op[k])
                gbs = by[i] / ms[i] / 1e6
                points.append({"c": c, "stride": stride, "batch": batch, "us": ms[i] * 1e3, "bytes": by[i],
                               "GBps": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"]})
                print(f"C={c:5d} s{stride}: {ms[i] * 1e3:8.1f} us  {gbs:7.0f} GB/s  {100 * gbs / peaks['hbm_gbs']:5.1f} %",
                      file=sys.stderr)
    print(json.dumps({"config": "depthwise_conv2d 3x3 int8 (+fused relu table), H=W=56, pad 1, batch %d, 1xB200" % batch,
                      "peak_hbm_gbs": peaks["hbm_gbs"], "points": points}))


if __name__ == "__main__":
    main()
