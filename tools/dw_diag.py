"""developer diagnostic: time the int8 depthwise 3x3 step of a one-layer session under the kernel's
diagnostic switches (SHL_B200_DW_UMMA / SHL_B200_DW_UMMA_DIAG); results are NOT checked"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from shl import DT_INT8, H_CONV, RM_GRAPH, Harness, Layer, synth_conv_i8  # noqa: E402

shl = C.CDLL(os.path.join(ROOT, "csi-nn2_b200", "lib", "libshl_b200.so"))
shl.shl_b200_session_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.c_int]
b200 = Harness("b200")
rng = np.random.default_rng(0)
for (n, c, hw) in [(256, 512, 14), (256, 128, 56), (256, 32, 112)]:
    x = rng.integers(-128, 128, size=(n, c, hw, hw), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
    layers = [Layer(H_CONV, (n, c, hw, hw), s_out=s_out, zp_out=0, w=wt, b=b, s_w=s_w, pad=(1,) * 4, group=c)]
    with b200.create(DT_INT8, x.shape, layers, s_in=0.02, zp_in=-128, run_mode=RM_GRAPH) as net:
        net(x)
        for env in sys.argv[1:] or ["SHL_B200_DW_UMMA=0", "SHL_B200_DW_UMMA=1"]:
            kv = dict(e.split("=") for e in env.split(","))
            os.environ.update(kv)
            ms, by, op = (C.c_double * 8)(), (C.c_double * 8)(), (C.c_double * 8)()
            k = shl.shl_b200_session_profile(net.session, 3, 10, ms, by, op, 8)
            i = max(range(k), key=lambda j: op[j])
            print(f"n{n} c{c} {hw}x{hw} {env:40s} {ms[i] * 1e3:8.1f} us {by[i] / ms[i] / 1e6:7.0f} GB/s", flush=True)
            for key in kv:
                os.environ.pop(key, None)
