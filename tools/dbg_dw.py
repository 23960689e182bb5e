"""developer diagnostic: where does the int8 depthwise 3x3 differ from the oracle? (GPU box)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from shl import DT_INT8, H_CONV, Harness, Layer, Oracle, synth_conv_i8

b200, oracle = Harness("b200"), Oracle()
rng = np.random.default_rng(0)
cases = [(2, 128, 13, 29, -7), (5, 256, 7, 7, 3), (2, 128, 56, 56, -128), (2, 32, 13, 29, -7), (1, 64, 56, 56, 0), (7, 48, 7, 7, 5), (1, 16, 7, 7, -128), (1, 512, 14, 14, -128),
         (3, 32, 112, 112, -128), (256, 32, 14, 14, 3)]
for (n, c, h, w, zp_in) in cases:
    x = rng.integers(-128, 128, size=(n, c, h, w), dtype=np.int8)
    wt, s_w, b, s_out = synth_conv_i8(rng, c, c, 3, 3, depthwise=True)
    layer = Layer(H_CONV, (n, c, h, w), s_out=s_out, zp_out=2, w=wt, b=b, s_w=s_w, pad=(1,) * 4, group=c)
    got = b200.run(DT_INT8, x.shape, [layer], x, s_in=0.02, zp_in=zp_in)
    want = oracle.conv2d_i8(x, wt, b, x.shape, depthwise=True, stride=(1, 1), pad=(1,) * 4, dilation=(1, 1), group=1,
                            s_in=0.02, zp_in=zp_in, s_w=s_w, s_b=None, s_out=s_out, zp_out=2)
    bad = np.argwhere(got != want)
    print(f"case n{n} c{c} {h}x{w}: {len(bad)}/{got.size} differ")
    if len(bad):
        for ax, nm in enumerate("ncyx"):
            vals, cnt = np.unique(bad[:, ax], return_counts=True)
            print("   ", nm, dict(zip(vals.tolist()[:40], cnt.tolist()[:40])))
        print("    first:", bad[:6].tolist())
