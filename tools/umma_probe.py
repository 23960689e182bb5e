"""developer probe (GPU box): SWIZZLE_128B A descriptors whose start is shifted by whole 128-byte rows
(csrc/umma_probe.cu) -- for which shifts does the MMA read the rows TMA wrote?"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shim = C.CDLL(os.path.join(ROOT, "csi-nn2_b200", "lib", "libb200nn.so"))
shim.b200_last_error.restype = C.c_char_p
rng = np.random.default_rng(0)
a = rng.integers(-128, 128, size=(144, 128), dtype=np.int8)
b = rng.integers(-128, 128, size=(32, 128), dtype=np.int8)
d_a, d_b, d_o = C.c_void_p(), C.c_void_p(), C.c_void_p()
for d, n in ((d_a, a.nbytes), (d_b, b.nbytes), (d_o, 128 * 32 * 4)):
    assert shim.b200_malloc(C.byref(d), C.c_size_t(n)) == 0, shim.b200_last_error()
assert shim.b200_memcpy_h2d(d_a, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), None) == 0
assert shim.b200_memcpy_h2d(d_b, b.ctypes.data_as(C.c_void_p), C.c_size_t(b.nbytes), None) == 0
ok = {}
for k0 in range(4):
    for shift in range(17):
        got = np.zeros((128, 32), np.int32)
        assert shim.b200_test_umma_shifted_start(d_a, d_b, shift, k0, d_o, None) == 0, shim.b200_last_error()
        assert shim.b200_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_o, C.c_size_t(got.nbytes), None) == 0
        assert shim.b200_stream_sync(None) == 0, shim.b200_last_error()
        ks = slice(32 * k0, 32 * k0 + 32)
        want = a[shift:shift + 128, ks].astype(np.int32) @ b[:, ks].astype(np.int32).T
        ok[(k0, shift)] = bool(np.array_equal(got, want))
        if not ok[(k0, shift)] and k0 == 0:
            # which input row does each output row look like?
            cand = a[:, ks].astype(np.int32) @ b[:, ks].astype(np.int32).T
            rows = [int(np.argmax((cand == got[m]).all(axis=1))) if (cand == got[m]).all(axis=1).any() else -1 for m in range(16)]
            print(f"k0 {k0} shift {shift}: first output rows read input rows {rows}")
for k0 in range(4):
    print("k0", k0, "shifts that match:", [s for s in range(17) if ok[(k0, s)]])

# ---- issue rate of small MMAs on one SM ----
shim.b200_test_umma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_void_p]
print("cycles per M128 x N x K32 kind::i8 MMA (one SM, 2000 back to back):")
for n in (16, 32, 64, 128, 256):
    row = []
    for nacc in (1, 2, 512 // n if 512 // n < 8 else 8):
        cyc = C.c_longlong()
        assert shim.b200_test_umma_rate(n, nacc, 2000, C.byref(cyc), None) == 0, shim.b200_last_error()
        row.append(f"nacc {nacc}: {cyc.value / 2000:6.1f}")
    print(f"  N {n:3d}   " + "   ".join(row))

shim.b200_test_umma_rate2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_void_p]
for m, f16, name in ((64, 0, "M64 i8"), (128, 1, "M128 f16 (K = 16 halves)"), (64, 1, "M64 f16")):
    row = []
    for n in (16, 64, 128, 256):
        cyc = C.c_longlong()
        assert shim.b200_test_umma_rate2(m, n, f16, 2, 2000, C.byref(cyc), None) == 0, shim.b200_last_error()
        row.append(f"N {n}: {cyc.value / 2000:6.1f}")
    print(f"  {name:26s}" + "   ".join(row))

shim.b200_test_umma_rate3.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_void_p]
print("2000 M128 i8 MMAs in total, issued by 1 / 2 / 4 warps (cycles per MMA):")
for n in (32, 128):
    row = []
    for issuers in (1, 2, 4):
        cyc = C.c_longlong()
        assert shim.b200_test_umma_rate3(128, n, 0, 4, 2000, issuers, C.byref(cyc), None) == 0, shim.b200_last_error()
        row.append(f"{issuers} warps: {cyc.value / 2000:6.1f}")
    print(f"  N {n:3d}   " + "   ".join(row))
