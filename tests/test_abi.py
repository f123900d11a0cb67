"""C-ABI surface checks that need no GPU: the library loads, exports exactly
what include/opengjk_b200.h declares, struct layouts match the reference's, and
the product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from opengjk_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for hdr in ("opengjk_b200.h",):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        names += re.findall(r"OGJK_API\s+[\w\s\*]+?\b(ogjk_\w+)\s*\(", src)
    return names


def test_header_and_binding_agree():
    assert sorted(set(_declared())) == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name


def test_version_string():
    assert b"sm_100a" in _lib.lib().ogjk_version()


def test_struct_layouts_match_reference():
    # fp32 gkSimplex 108 B {0,4,52,84}; fp64 184 B {0,8,104,136}
    f32, f64 = _lib.SIMPLEX_F32, _lib.SIMPLEX_F64
    assert [f32.fields[k][1] for k in ("nvrtx", "vrtx", "vrtx_idx", "witnesses")] == [0, 4, 52, 84]
    assert [f64.fields[k][1] for k in ("nvrtx", "vrtx", "vrtx_idx", "witnesses")] == [0, 8, 104, 136]
    P32 = api._polytope_struct(C.c_float)
    P64 = api._polytope_struct(C.c_double)
    assert C.sizeof(P32) == 32 and C.sizeof(P64) == 48
    assert P32.coord.offset == 24 and P64.coord.offset == 40 and P64.s_idx.offset == 32


def test_no_cpu_fallback_without_device(has_cuda):
    if has_cuda:
        pytest.skip("CUDA device present")
    with pytest.raises(api.OpenGJKError, match="no CUDA device|CUDA"):
        api.Engine(0)


def test_argument_errors_do_not_need_a_device():
    L = _lib.lib()
    assert L.ogjk_ctx_create(0, None) == -1
    assert b"NULL" in L.ogjk_last_error()
    z = np.zeros(1, dtype=np.float32)
    rc = L.ogjk_gjk_device_f32(None, 1, z.ctypes.data, None, None, 1, 1, None, None, None, None, 1, 1, None, None, None, 0,
                                None)
    assert rc == -1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opengjk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "gjk_oracle" not in text and "oracle/" not in text, f


def test_tensor_front_door_rejects_host_tensors():
    """opengjk_b200.torch_api is device-only: host tensors are refused before any engine call."""
    import pytest
    import torch
    from opengjk_b200 import torch_api
    with pytest.raises(ValueError, match="CUDA tensors required"):
        torch_api.compute_minimum_distance(torch.zeros(2, 4, 3), torch.zeros(2, 4, 3))
