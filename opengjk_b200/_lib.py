"""ctypes binding of the C ABI in include/opengjk_b200.h.

The shared library is built in-tree (opengjk_b200/libopengjk_b200.so) by
``make -C opengjk_b200/csrc`` / ``__graft_entry__.build()``.  There is no
fallback of any kind: a missing library or a missing CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libopengjk_b200.so")
# Same ABI plus the measured-and-lost kernel flavours (-DOGJK_EXPERIMENTAL); tests only.
EXP_LIB_PATH = os.path.join(_HERE, "libopengjk_b200_exp.so")

_libs = {}


class OpenGJKError(RuntimeError):
    pass


def simplex_dtype(np_float):
    """numpy mirror of gkSimplex (scalar/include/openGJK/openGJK.h:84-94)."""
    return np.dtype([("nvrtx", np.int32), ("vrtx", np_float, (4, 3)), ("vrtx_idx", np.int32, (4, 2)),
                     ("witnesses", np_float, (2, 3))], align=True)


SIMPLEX_F32 = simplex_dtype(np.float32)
SIMPLEX_F64 = simplex_dtype(np.float64)
assert SIMPLEX_F32.itemsize == 108 and SIMPLEX_F64.itemsize == 184

# Every symbol include/opengjk_b200.h declares (tests check the .so exports all of them).
EXPORTS = [
    "ogjk_last_error", "ogjk_version", "ogjk_ctx_create", "ogjk_ctx_destroy", "ogjk_ctx_device",
    "ogjk_ctx_launch_count", "ogjk_ctx_pool_status", "ogjk_ctx_set_warp_threshold", "ogjk_ctx_set_timing", "ogjk_ctx_last_timing",
    "ogjk_ctx_set_classes", "ogjk_ctx_set_pipeline_chunks", "ogjk_ctx_set_accel_min", "ogjk_ctx_set_table_lanes", "ogjk_ctx_set_regroup", "ogjk_ctx_set_passes", "ogjk_store_create_f32", "ogjk_store_create_f64", "ogjk_store_destroy", "ogjk_store_bytes",
    "ogjk_store_num_polytopes", "ogjk_store_table_bytes", "ogjk_store_aabbs_f32", "ogjk_store_aabbs_f64", "ogjk_pairs_cull_f32",
    "ogjk_pairs_cull_f64", "ogjk_pairs_all_f32", "ogjk_pairs_all_f64", "ogjk_intersect_flags_f32", "ogjk_intersect_flags_f64", "ogjk_store_gjk_f32", "ogjk_store_gjk_f64",
    "ogjk_store_epa_f32", "ogjk_store_epa_f64", "ogjk_store_gjk_epa_f32", "ogjk_store_gjk_epa_f64",
    "ogjk_gjk_epa_host_f32", "ogjk_gjk_epa_host_f64", "ogjk_compute_gjk_epa_f32", "ogjk_compute_gjk_epa_f64",
    "ogjk_compute_epa_f32", "ogjk_compute_epa_f64", "ogjk_test_subalgorithm_f32", "ogjk_test_subalgorithm_f64",
    "ogjk_gjk_device_structs_f32", "ogjk_gjk_device_structs_f64", "ogjk_epa_device_structs_f32",
    "ogjk_epa_device_structs_f64",
    "ogjk_gjk_device_f32", "ogjk_gjk_device_f64", "ogjk_gjk_host_f32", "ogjk_gjk_host_f64",
    "ogjk_compute_minimum_distance_f32", "ogjk_compute_minimum_distance_f64",
    "ogjk_compute_minimum_distance_indexed_f32", "ogjk_compute_minimum_distance_indexed_f64",
    "ogjk_single_rows_f32", "ogjk_single_rows_f64",
    # several GPUs behind one call (csrc/multi.cuh)
    "ogjk_multi_allgather_f32",
    "ogjk_multi_compute_gjk_epa_f32",
    "ogjk_multi_compute_gjk_epa_f64",
    "ogjk_multi_compute_minimum_distance_f32",
    "ogjk_multi_compute_minimum_distance_f64",
    "ogjk_multi_allgather_f64",
    "ogjk_multi_create",
    "ogjk_multi_ctx",
    "ogjk_multi_destroy",
    "ogjk_multi_gjk_epa_host_f32",
    "ogjk_multi_gjk_epa_host_f64",
    "ogjk_multi_gjk_host_f32",
    "ogjk_multi_gjk_host_f64",
    "ogjk_multi_launch_count",
    "ogjk_multi_num_devices",
    "ogjk_multi_resident_f32",
    "ogjk_multi_resident_f64",
    "ogjk_multi_shard_bounds",
    "ogjk_multi_store_bytes",
    "ogjk_multi_store_create_f32",
    "ogjk_multi_store_create_f64",
    "ogjk_multi_store_destroy",
    "ogjk_multi_store_gjk_epa_f32",
    "ogjk_multi_store_gjk_epa_f64",
    "ogjk_multi_store_gjk_f32",
    "ogjk_multi_store_gjk_f64",
    "ogjk_multi_store_gjk_resident_f32",
    "ogjk_multi_store_gjk_resident_f64",
]


def lib(path=None):
    """The product library (default) or another build of the same ABI, loaded once per path."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise OpenGJKError(
            f"{path} is missing: build it with `make -C opengjk_b200/csrc` "
            "(there is no CPU fallback)")
    L = C.CDLL(path)
    L.ogjk_last_error.restype = C.c_char_p
    L.ogjk_version.restype = C.c_char_p
    L.ogjk_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ogjk_ctx_destroy.argtypes = [C.c_void_p]
    L.ogjk_ctx_destroy.restype = None
    L.ogjk_ctx_device.argtypes = [C.c_void_p]
    L.ogjk_ctx_launch_count.argtypes = [C.c_void_p]
    L.ogjk_ctx_launch_count.restype = C.c_int64
    L.ogjk_ctx_set_warp_threshold.argtypes = [C.c_void_p, C.c_int]
    L.ogjk_ctx_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.ogjk_ctx_last_timing.argtypes = [C.c_void_p, C.c_void_p]
    L.ogjk_ctx_set_classes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ogjk_ctx_set_pipeline_chunks.argtypes = [C.c_void_p, C.c_int]
    L.ogjk_ctx_set_accel_min.argtypes = [C.c_void_p, C.c_int]
    L.ogjk_ctx_set_table_lanes.argtypes = [C.c_void_p, C.c_int]
    L.ogjk_ctx_set_regroup.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ogjk_ctx_set_passes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ogjk_store_destroy.argtypes = [C.c_void_p]
    L.ogjk_store_destroy.restype = None
    L.ogjk_store_bytes.argtypes = [C.c_void_p]
    L.ogjk_store_bytes.restype = C.c_int64
    L.ogjk_store_num_polytopes.argtypes = [C.c_void_p]
    L.ogjk_store_num_polytopes.restype = C.c_int64
    for sfx, ct in (("f32", C.c_float), ("f64", C.c_double)):
        getattr(L, f"ogjk_store_aabbs_{sfx}").argtypes = [C.c_void_p] * 4
        getattr(L, f"ogjk_pairs_cull_{sfx}").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                         ct, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        getattr(L, f"ogjk_pairs_all_{sfx}").argtypes = [C.c_void_p, C.c_void_p, C.c_int64, ct, C.c_int64, C.c_void_p,
                                                        C.c_void_p, C.c_void_p, C.c_void_p]
    L.ogjk_intersect_flags_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ogjk_intersect_flags_f64.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ogjk_store_table_bytes.argtypes = [C.c_void_p]
    L.ogjk_store_table_bytes.restype = C.c_int64
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    for sfx in ("f32", "f64"):
        getattr(L, f"ogjk_gjk_device_{sfx}").argtypes = [vp, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, i64, i64, vp,
                                                         vp, vp, i32, vp]
        getattr(L, f"ogjk_store_create_{sfx}").argtypes = [vp, i64, i64, vp, vp, vp, i32, vp, C.POINTER(vp)]
        getattr(L, f"ogjk_store_gjk_{sfx}").argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp]
        getattr(L, f"ogjk_store_epa_{sfx}").argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_store_gjk_epa_{sfx}").argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_gjk_epa_host_{sfx}").argtypes = [vp, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, i64, i64, vp,
                                                           vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_compute_gjk_epa_{sfx}").argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_compute_epa_{sfx}").argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_test_subalgorithm_{sfx}").argtypes = [vp, vp, vp]
        getattr(L, f"ogjk_gjk_device_structs_{sfx}").argtypes = [vp, i64, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_epa_device_structs_{sfx}").argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_gjk_host_{sfx}").argtypes = [vp, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, i64, i64, vp,
                                                       vp, vp]
        getattr(L, f"ogjk_compute_minimum_distance_{sfx}").argtypes = [vp, i32, vp, vp, vp, vp]
        getattr(L, f"ogjk_compute_minimum_distance_indexed_{sfx}").argtypes = [vp, i32, i32, vp, vp, vp, vp]
        getattr(L, f"ogjk_single_rows_{sfx}").argtypes = [vp, i32, vp, i32, vp, vp, vp]
    L.ogjk_ctx_pool_status.argtypes = [C.c_void_p]
    L.ogjk_multi_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.ogjk_multi_destroy.argtypes = [vp]
    L.ogjk_multi_destroy.restype = None
    L.ogjk_multi_num_devices.argtypes = [vp]
    L.ogjk_multi_ctx.argtypes = [vp, i32]
    L.ogjk_multi_ctx.restype = vp
    L.ogjk_multi_launch_count.argtypes = [vp]
    L.ogjk_multi_launch_count.restype = i64
    L.ogjk_multi_shard_bounds.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    L.ogjk_multi_shard_bounds.restype = None
    L.ogjk_multi_store_destroy.argtypes = [vp]
    L.ogjk_multi_store_destroy.restype = None
    L.ogjk_multi_store_bytes.argtypes = [vp]
    L.ogjk_multi_store_bytes.restype = i64
    for sfx in ("f32", "f64"):
        getattr(L, f"ogjk_multi_gjk_host_{sfx}").argtypes = [vp, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, i64, i64, vp,
                                                             vp, vp]
        getattr(L, f"ogjk_multi_gjk_epa_host_{sfx}").argtypes = [vp, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, i64, i64,
                                                                 vp, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_multi_compute_minimum_distance_{sfx}").argtypes = [vp, i32, vp, vp, vp, vp]
        getattr(L, f"ogjk_multi_compute_gjk_epa_{sfx}").argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_multi_store_create_{sfx}").argtypes = [vp, i64, i64, vp, vp, vp, i32, C.POINTER(vp)]
        getattr(L, f"ogjk_multi_store_gjk_{sfx}").argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
        getattr(L, f"ogjk_multi_store_gjk_epa_{sfx}").argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
        getattr(L, f"ogjk_multi_store_gjk_resident_{sfx}").argtypes = [vp, vp, vp, vp, vp, i64]
        getattr(L, f"ogjk_multi_allgather_{sfx}").argtypes = [vp]
        getattr(L, f"ogjk_multi_resident_{sfx}").argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64),
                                                             C.POINTER(i64)]
    _libs[path] = L
    return L


def check(rc, L=None):
    if rc != 0:
        raise OpenGJKError(f"opengjk_b200 error {rc}: {(L or lib()).ogjk_last_error().decode()}")
