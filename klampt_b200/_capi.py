"""ctypes binding of the C ABI in include/klampt_b200.h (libklampt_b200.so, built in-tree by
``__graft_entry__.build()``).  No torch types cross this boundary: plain pointers and sizes only.

The library is the product; if it is missing or does not load, importing the engine fails loudly --
there is no Python / CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KLAMPT_B200_LIB", os.path.join(_HERE, "libklampt_b200.so"))

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_int64_p = C.POINTER(C.c_int64)


class KbStats(C.Structure):
    _fields_ = [("configs_checked", C.c_int64), ("configs_feasible", C.c_int64), ("edges_checked", C.c_int64),
                ("edges_visible", C.c_int64), ("edge_config_checks", C.c_int64), ("node_tests", C.c_int64),
                ("elem_tests", C.c_int64), ("recheck_pairs", C.c_int64), ("kernel_launches", C.c_int64),
                ("traverse_launches", C.c_int64), ("traverse_ms", C.c_double), ("gpu_ms", C.c_double),
                ("items_dropped", C.c_int64), ("node_iterations", C.c_int64), ("rays_cast", C.c_int64)]


class KbCamera(C.Structure):
    _fields_ = [("pose", C.c_double * 12), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("zmin", C.c_double), ("zmax", C.c_double), ("xres", C.c_int32), ("yres", C.c_int32)]


# every symbol include/klampt_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SIGNATURES = {
    "kb_engine_create": (C.c_int, [C.POINTER(_VP)]),
    "kb_engine_destroy": (None, [_VP]),
    "kb_last_error": (C.c_char_p, []),
    "kb_version": (C.c_char_p, []),
    "kb_add_trimesh": (C.c_int, [_VP, c_double_p, C.c_int, c_int32_p, C.c_int, C.c_double]),
    "kb_add_pointcloud": (C.c_int, [_VP, c_double_p, C.c_int, c_double_p, C.c_double]),
    "kb_add_primitive": (C.c_int, [_VP, C.c_int, c_double_p, C.c_double]),
    "kb_add_dynamic_pointcloud": (C.c_int, [_VP, C.c_int, C.c_double, C.c_double]),
    "kb_update_pointcloud": (C.c_int, [_VP, C.c_int, _VP, C.c_int]),
    "kb_add_terrain": (C.c_int, [_VP, C.c_int]),
    "kb_add_rigid_object": (C.c_int, [_VP, C.c_int, c_double_p]),
    "kb_robot_create": (C.c_int, [_VP, C.c_int, c_int32_p, c_uint8_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "kb_robot_set_link_geometry": (C.c_int, [_VP, C.c_int, C.c_int]),
    "kb_robot_set_joints": (C.c_int, [_VP, C.c_int, c_uint8_p, c_int32_p, c_int32_p]),
    "kb_robot_add_driver": (C.c_int, [_VP, C.c_int, c_int32_p, c_double_p, c_double_p, C.c_double, C.c_double]),
    "kb_robot_set_self_collision": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int]),
    "kb_set_pair_mask": (C.c_int, [_VP, c_uint8_p, C.c_int]),
    "kb_num_ids": (C.c_int, [_VP]),
    "kb_get_pair_mask": (C.c_int, [_VP, c_uint8_p]),
    "kb_finalize": (C.c_int, [_VP, C.c_int]),
    "kb_finalize_multi": (C.c_int, [_VP, c_int32_p, C.c_int]),
    "kb_num_devices": (C.c_int, [_VP]),
    "kb_set_stream": (C.c_int, [_VP, _VP]),
    "kb_set_option": (C.c_int, [_VP, C.c_char_p, C.c_int64]),
    "kb_synchronize": (C.c_int, [_VP]),
    "kb_fk_batch": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "kb_feasible_batch": (C.c_int, [_VP, _VP, C.c_int64, _VP, _VP]),
    "kb_feasible_batch_f32": (C.c_int, [_VP, _VP, C.c_int64, _VP, _VP]),
    "kb_feasible_batch_device": (C.c_int, [_VP, _VP, C.c_int64, _VP, _VP]),
    "kb_feasible_batch_bits": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "kb_feasible_batch_bits_device": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "kb_edges_visible_batch_bits": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_double, _VP, _VP, _VP]),
    "kb_edges_visible_batch": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_double, _VP, _VP, _VP]),
    "kb_edges_visible_batch_device": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_double, _VP, _VP, _VP]),
    "kb_colliding_pairs_batch": (C.c_int, [_VP, _VP, C.c_int64, C.c_int, _VP, _VP]),
    "kb_distance_batch": (C.c_int, [_VP, _VP, C.c_int64, C.c_double, C.c_int, _VP, _VP]),
    "kb_distance_batch_device": (C.c_int, [_VP, _VP, C.c_int64, C.c_double, C.c_int, _VP, _VP]),
    "kb_distance_batch_ex": (C.c_int, [_VP, _VP, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int, _VP, _VP, _VP, _VP]),
    "kb_geom_distance_batch_ex": (C.c_int, [_VP, C.c_int, _VP, C.c_int, _VP, C.c_int64, C.c_double, C.c_double, C.c_double, _VP, _VP, _VP]),
    "kb_raycast_batch": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP]),
    "kb_raycast_batch_f32": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP]),
    "kb_raycast_batch_device": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP]),
    "kb_geom_raycast_batch": (C.c_int, [_VP, C.c_int, _VP, _VP, C.c_int64, _VP, _VP]),
    "kb_camera_depth": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "kb_geom_collides_batch": (C.c_int, [_VP, C.c_int, _VP, C.c_int, _VP, C.c_int64, C.c_double, _VP]),
    "kb_geom_distance_batch": (C.c_int, [_VP, C.c_int, _VP, C.c_int, _VP, C.c_int64, C.c_double, _VP]),
    "kb_get_stats": (C.c_int, [_VP, C.POINTER(KbStats)]),
    "kb_reset_stats": (C.c_int, [_VP]),
    "kb_get_layout": (C.c_int, [_VP, c_int64_p]),
}

_lib = None


class KbError(RuntimeError):
    pass


def load():
    """Loads libklampt_b200.so and types every entry point.  Raises ImportError if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("klampt_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        raise KbError("klampt_b200 error %d: %s" % (rc, load().kb_last_error().decode("utf-8", "replace")))
    return rc
