"""ctypes binding of ``libqsmrt.so`` (C ABI declared in ``include/qsmrt.h``).

The library is built in-tree by ``pyqsm_b200/csrc/Makefile`` (nvcc, sm_100a).
There is no CPU fallback: if the library is missing or no CUDA device is
present, loading / scene creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QSMRT_LIB") or os.path.join(_HERE, "libqsmrt.so")      # QSMRT_LIB: an A/B build of the same ABI
ABI_VERSION = 2

_lib = None


class Stats(C.Structure):
    _fields_ = [
        ("num_triangles", C.c_uint64), ("num_geometries", C.c_uint64),
        ("num_bvh_nodes", C.c_uint64), ("num_bvh_leaves", C.c_uint64),
        ("bvh_bytes", C.c_uint64),
        ("build_ms", C.c_float), ("sort_ms", C.c_float), ("box_pad", C.c_float),
        ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3),
        ("leaf_max", C.c_uint32), ("bvh_height", C.c_uint32), ("quantised_nodes", C.c_uint32), ("full_sort", C.c_uint32),
        ("num_references", C.c_uint64),
    ]


# enum qsmrt_option (include/qsmrt.h)
OPT = {
    "leaf_max": 1, "keep_binary_nodes": 2, "quant_threshold": 3, "climb_capacity": 4, "sort_variant": 5, "split_max": 6, "split_aspect": 7,
    "quantised_nodes": 16, "traversal_variant": 17, "refill": 18, "want": 19, "tri_min": 20, "counters": 21,
    "node_path": 22, "cp_warp_max": 23, "ctas_per_sm": 24, "count_set": 25, "tile_order": 26, "point_order": 27, "host_chunk": 28, "host_ramp": 29,
}
SAVE_BVH = 1

# name -> (restype, argtypes); every symbol include/qsmrt.h declares
_vp, _u64, _f = C.c_void_p, C.c_uint64, C.c_float
SYMBOLS = {
    "qsmrt_last_error": (C.c_char_p, []),
    "qsmrt_abi_version": (C.c_int, []),
    "qsmrt_scene_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "qsmrt_scene_destroy": (C.c_int, [_vp]),
    "qsmrt_add_triangles": (C.c_int, [_vp, _vp, _u64, _vp, _u64, C.c_int, C.POINTER(C.c_uint32)]),
    "qsmrt_add_cylinders": (C.c_int, [_vp, _vp, _u64, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]),
    "qsmrt_geometry_size": (C.c_int, [_vp, C.c_uint32, C.POINTER(_u64), C.POINTER(_u64)]),
    "qsmrt_copy_geometry": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp]),
    "qsmrt_commit": (C.c_int, [_vp, _vp, C.POINTER(_f)]),
    "qsmrt_cast_rays": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qsmrt_cast_rays_2d": (C.c_int, [_vp, _vp, C.c_uint32, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qsmrt_cast_rays_host": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "qsmrt_count_intersections": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "qsmrt_test_occlusions": (C.c_int, [_vp, _vp, _u64, _f, _f, _vp, _vp]),
    "qsmrt_list_intersections_count": (C.c_int, [_vp, _vp, _u64, _vp, C.POINTER(C.c_int64), _vp]),
    "qsmrt_list_intersections_fill": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qsmrt_gen_parallel_rays": (C.c_int, [_vp, _u64, _u64, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), _vp]),
    "qsmrt_gen_pinhole_rays": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp]),
    "qsmrt_mark_hit_primitives": (C.c_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "qsmrt_accumulate_hits": (C.c_int, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "qsmrt_closest_points": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qsmrt_signed_distance": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "qsmrt_sun_exposure": (C.c_int, [_vp, _u64, _u64, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), _vp, _vp]),
    "qsmrt_sun_exposure_sweep": (C.c_int, [_vp, C.c_uint32, C.POINTER(_f), _u64, _u64, _vp, _u64, _vp]),
    "qsmrt_sky_visibility": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _u64, _f, C.c_uint32, C.c_uint32, _vp, _vp]),
    "qsmrt_gen_hemisphere_rays": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _u64, _f, C.c_uint32, C.c_uint32, _vp]),
    "qsmrt_peel_projection": (C.c_int, [_vp, _u64, _u64, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), C.c_int, _vp,
                                        C.POINTER(C.c_double), C.POINTER(C.c_int), _vp]),
    "qsmrt_util_read_sweep": (C.c_int, [_vp, _u64, C.c_uint32, _vp, _vp]),
    "qsmrt_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "qsmrt_debug_get_build": (C.c_int, [_vp, _vp, _vp, _vp]),
    "qsmrt_release_cached_memory": (C.c_int, []),
    "qsmrt_scene_set_option": (C.c_int, [_vp, C.c_int, C.c_double]),
    "qsmrt_scene_get_option": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "qsmrt_scene_get_counters": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "qsmrt_scene_save": (C.c_int, [_vp, C.c_char_p, C.c_uint32]),
    "qsmrt_scene_load": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(_vp)]),
    "qsmrt_cast_rays_host_split": (C.c_int, [_vp, _vp, _u64, C.POINTER(_vp), C.POINTER(_vp)]),
    "qsmrt_count_intersections_host": (C.c_int, [_vp, _vp, _u64, _vp]),
    "qsmrt_test_occlusions_host": (C.c_int, [_vp, _vp, _u64, _f, _f, _vp]),
    "qsmrt_vertex_exposure": (C.c_int, [_vp, _vp, _vp, _vp]),
}


def load():
    """Load libqsmrt.so and bind every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C pyqsm_b200/csrc` "
            "(or __graft_entry__.build()).  pyqsm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.qsmrt_abi_version() != ABI_VERSION:
        raise RuntimeError("libqsmrt.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().qsmrt_last_error()
        raise RuntimeError("qsmrt: " + (msg.decode() if msg else "unknown error"))
