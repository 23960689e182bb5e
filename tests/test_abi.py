"""CPU-side checks of the drop-in boundary: the shared libraries load without a GPU and export
every symbol the headers under include/ declare, the backend registers itself under the api ids
it claims, and -- with no device -- every compute entry fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from shl import (API_C906, API_C908, API_C920, API_C920V2, API_RVV, DT_F16, DT_F32, DT_INT8, H_CONV, ROOT,
                 Harness, Layer)

LIB = os.path.join(ROOT, "csi-nn2_b200", "lib")
DECL = re.compile(r"^\s*(?:const\s+)?(?:struct\s+\w+|unsigned\s+\w+|\w+)\s*\**\s*\b((?:b200|shl)_\w+)\s*\(", re.M)


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = []
    for m in DECL.finditer(text):
        line_start = text.rfind("\n", 0, m.start()) + 1
        if "static" in text[line_start:m.start(1)]:
            continue
        names.append(m.group(1))
    return sorted(set(names))


def test_shim_exports_every_declared_symbol():
    lib = C.CDLL(os.path.join(LIB, "libb200nn.so"))
    names = declared("b200nn.h")
    assert len(names) >= 35, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.b200_abi_version() == 3


def test_backend_exports_every_declared_symbol():
    lib = C.CDLL(os.path.join(LIB, "libshl_b200.so"))
    names = declared("shl_b200.h")
    assert len(names) >= 40, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the reference's public API is what the library serves
    for n in ("csinn_alloc_session", "csinn_conv2d_init", "csinn_conv2d", "csinn_depthwise_conv2d",
              "csinn_fullyconnected", "csinn_session_init", "csinn_session_setup", "csinn_session_run",
              "csinn_update_input", "csinn_get_output", "shl_register_op_callback"):
        assert hasattr(lib, n), n


def test_no_reference_operator_backend_is_linked():
    """No CPU compute path hides in the product: the reference's operator kernels
    (shl_ref_conv2d_quant, ...) must not be present in libshl_b200.so."""
    lib = C.CDLL(os.path.join(LIB, "libshl_b200.so"))
    for n in ("shl_ref_conv2d_quant", "shl_ref_conv2d_f32", "shl_ref_depthwise_conv2d_quant",
              "shl_ref_fullyconnected_quant", "shl_ref_relu_quant", "shl_target_init_ref", "shl_cb_map_ref"):
        assert not hasattr(lib, n), f"{n} is linked into the product"


def test_registered_under_rvv_and_c9xx_ids():
    lib = C.CDLL(os.path.join(LIB, "libshl_b200.so"))
    lib.csinn_alloc_session.restype = C.c_void_p
    lib.csinn_alloc_session()  # triggers shl_init() -> shl_target_init_rvv/c906/...
    lib.shl_cb_map_b200.restype = C.c_void_p

    class Callback(C.Structure):
        _fields_ = [(k, C.c_void_p) for k in ("init", "est", "exec", "caps", "perf")]

    CSINN_OP_CONV2D, CSINN_OP_ABS = None, 0
    # enum value of CSINN_OP_CONV2D: ask the table for every op until conv2d_init shows up
    init_addr = C.cast(lib.shl_b200_conv2d_init, C.c_void_p).value
    hits = []
    for op in range(0, 400):
        cb = C.cast(lib.shl_cb_map_b200(op, DT_INT8), C.POINTER(Callback)).contents
        if cb.init == init_addr:
            hits.append(op)
    assert len(hits) == 2, hits  # CONV2D and GROUP_CONV2D share the init
    cb = C.cast(lib.shl_cb_map_b200(hits[0], DT_F16), C.POINTER(Callback)).contents
    assert cb.init == init_addr and cb.exec and cb.est
    # unsupported dtype -> zeroed callback, never a CPU fallback
    cb = C.cast(lib.shl_cb_map_b200(hits[0], DT_F32), C.POINTER(Callback)).contents
    assert not cb.init and not cb.exec and not cb.est


@pytest.mark.parametrize("api", [API_RVV, API_C906, API_C908, API_C920, API_C920V2])
def test_fails_loudly_without_device(api):
    """On a box without a GPU the product must refuse, not compute on the CPU."""
    lib = C.CDLL(os.path.join(LIB, "libb200nn.so"))
    if lib.b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = Harness("b200")
    w = np.ones((16, 16, 1, 1), np.int8)
    layer = Layer(H_CONV, (1, 16, 4, 4), s_out=1.0, w=w, b=np.zeros(16, np.int32), s_w=np.ones(16, np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        h.run(DT_INT8, (1, 16, 4, 4), [layer], np.zeros((1, 16, 4, 4), np.int8), api=api)


def test_requant_lut_matches_reference_formula():
    """b200_build_requant_lut (host) == the reference's dequant -> relu -> quant float sequence."""
    lib = C.CDLL(os.path.join(LIB, "libb200nn.so"))
    from shl import Oracle
    orc = Oracle()
    q = np.arange(-128, 128, dtype=np.int8)
    for act, s_in, zp_in, s_out, zp_out in [(1, 0.037, -3, 0.0181, -128), (2, 0.11, 5, 0.02352941, -128),
                                            (0, 0.05, 0, 0.05, 0), (1, 0.02, -128, 0.031, -7)]:
        lut = (C.c_int8 * 256)()
        lib.b200_build_requant_lut(lut, act, C.c_float(s_in), zp_in, C.c_float(s_out), zp_out)
        want = orc.relu_i8(q, act, s_in, zp_in, s_out, zp_out)
        assert np.array_equal(np.frombuffer(lut, np.int8), want)
