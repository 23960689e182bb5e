import sys, numpy as np
sys.path.insert(0, 'tests')
from shl import *
import oracle_np as onp
b200 = Harness("b200")
rng = np.random.default_rng(1234)
for sa, tb in [((4, 24, 24), False), ((1, 130, 64), True), ((1, 24, 24), False), ((1, 24, 24), True), ((2, 24, 24), True), ((1,128,64),True), ((1,129,64),True), ((1,130,32),True)]:
    a = rng.standard_normal(sa).astype(np.float16)
    j = sa[-2] if tb else sa[-1]
    out_shape = sa[:-2] + (sa[-2], j)
    layers = [Layer(H_RELU, sa, in0=0), Layer(H_MATMUL, out_shape, in0=0, in1=1, pad=(0, 0, 0, int(tb)))]
    got = b200.run(DT_F16, sa, layers, a, run_mode=RM_GRAPH).astype(np.float32)
    want = onp.matmul_f32(a, np.maximum(a, 0), False, tb)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1)
    bad = np.argwhere(err > 1e-3)
    print(sa, tb, "bad", len(bad), "of", err.size, "first", bad[:5].tolist(), "rows", sorted(set(bad[:, -2].tolist()))[:10], "cols", sorted(set(bad[:, -1].tolist()))[:10])
