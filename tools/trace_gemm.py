import os, sys
sys.path.insert(0, "tests")
import numpy as np
from shl import *
os.environ["SHL_B200_GEMM_TRACE"] = "1"
rng = np.random.default_rng(0)
b200 = Harness("b200")
from test_gpu_parity import synth_conv_i8
for (n, c, h, w, o) in [(256, 512, 14, 14, 512), (256, 128, 56, 56, 128), (256, 32, 112, 112, 64)]:
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, o, 1, 1)
    layer = Layer(H_CONV, (n, o, h, w), s_out=s_out, zp_out=3, w=wt, b=b, s_w=s_w)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, zp_in=0)
    print("done", n, c, h, w, o, got.shape)
