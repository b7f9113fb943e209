"""The C-ABI library loads and exports every symbol include/klampt_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

from klampt_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "klampt_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_hot_path():
    syms = header_symbols()
    for must in ("kb_engine_create", "kb_finalize", "kb_fk_batch", "kb_feasible_batch", "kb_feasible_batch_device",
                 "kb_edges_visible_batch", "kb_distance_batch", "kb_set_pair_mask", "kb_last_error", "kb_get_stats"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), "libklampt_b200.so does not export %s" % s


def test_python_binding_covers_the_header(built):
    assert sorted(_capi.SIGNATURES) == header_symbols()
    lib = _capi.load()
    assert b"sm_100a" in lib.kb_version()


def test_library_is_sm100a_native(built):
    """the fat binary carries sm_100a SASS (cuobjdump lists the ELF)"""
    import shutil
    import subprocess
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cu):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cu, "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_error_path_without_gpu_is_loud(built):
    """no CUDA device -> kb_finalize fails with KB_ERR_CUDA and a message; nothing falls back to the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from klampt_b200 import synth
    from klampt_b200.engine import Engine
    with pytest.raises(_capi.KbError) as ei:
        Engine(synth.world_c1())
    assert "no CPU fallback" in str(ei.value)


def test_argument_validation_without_gpu(built):
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.kb_engine_create(ctypes.byref(h)) == 0
    import numpy as np
    par = np.array([-1, 1], dtype=np.int32)          # parents[1] must be < 1
    z8 = np.zeros(2, dtype=np.uint8)
    zd = np.zeros(24)
    rc = lib.kb_robot_create(h, 2, par.ctypes.data_as(_capi.c_int32_p), z8.ctypes.data_as(_capi.c_uint8_p), zd.ctypes.data_as(_capi.c_double_p),
                             zd.ctypes.data_as(_capi.c_double_p), zd.ctypes.data_as(_capi.c_double_p), zd.ctypes.data_as(_capi.c_double_p))
    assert rc == -1 and b"parents" in lib.kb_last_error()
    assert lib.kb_feasible_batch(h, None, 1, None, None) == -2       # not finalized
    v = np.zeros((3, 3)); t = np.array([[0, 1, 5]], dtype=np.int32)  # vertex index out of range
    assert lib.kb_add_trimesh(h, v.ctypes.data_as(_capi.c_double_p), 3, t.ctypes.data_as(_capi.c_int32_p), 1, 0.0) == -1
    assert lib.kb_add_primitive(h, 7, zd.ctypes.data_as(_capi.c_double_p), 0.0) == -4
    lib.kb_engine_destroy(h)


def test_ingestion_rejects_non_finite_and_null_input_without_gpu(built):
    """scene ingestion is host code: a NaN coordinate, a null array or a negative radius is refused with KB_ERR_INVALID and a
    message instead of reaching the hierarchy builder"""
    import numpy as np
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.kb_engine_create(ctypes.byref(h)) == 0
    dp, ip = _capi.c_double_p, _capi.c_int32_p
    v = np.array([[0, 0, 0], [1, 0, 0], [0, np.nan, 0]], dtype=np.float64); t = np.array([[0, 1, 2]], dtype=np.int32)
    assert lib.kb_add_trimesh(h, v.ctypes.data_as(dp), 3, t.ctypes.data_as(ip), 1, 0.0) == -1 and b"non-finite" in lib.kb_last_error()
    assert lib.kb_add_trimesh(h, None, 3, t.ctypes.data_as(ip), 1, 0.0) == -1 and b"null" in lib.kb_last_error()
    v[2, 1] = 1.0
    assert lib.kb_add_trimesh(h, v.ctypes.data_as(dp), 3, t.ctypes.data_as(ip), 1, float("nan")) == -1
    g0 = lib.kb_add_trimesh(h, v.ctypes.data_as(dp), 3, t.ctypes.data_as(ip), 1, 0.0)
    assert g0 == 0
    p = np.array([[0, 0, 0], [np.inf, 0, 0]], dtype=np.float64); r = np.array([0.1, -0.1])
    assert lib.kb_add_pointcloud(h, p.ctypes.data_as(dp), 2, None, 0.0) == -1 and b"non-finite" in lib.kb_last_error()
    p[1, 0] = 1.0
    assert lib.kb_add_pointcloud(h, p.ctypes.data_as(dp), 2, r.ctypes.data_as(dp), 0.0) == -1 and b"negative radius" in lib.kb_last_error()
    assert lib.kb_add_pointcloud(h, None, 2, None, 0.0) == -1
    s = np.array([0, 0, 0, -1.0])
    assert lib.kb_add_primitive(h, 1, s.ctypes.data_as(dp), 0.0) == -1 and b"radius" in lib.kb_last_error()
    seg = np.array([1.0, 2, 3, 1, 2, 3])
    assert lib.kb_add_primitive(h, 5, seg.ctypes.data_as(dp), 0.0) == -1 and b"zero length" in lib.kb_last_error()
    seg[3] = np.nan
    assert lib.kb_add_primitive(h, 5, seg.ctypes.data_as(dp), 0.0) == -1 and b"non-finite" in lib.kb_last_error()
    seg[3] = 2.0
    assert lib.kb_add_primitive(h, 5, seg.ctypes.data_as(dp), 0.0) == 1
    T = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, np.nan])
    assert lib.kb_add_rigid_object(h, 0, T.ctypes.data_as(dp)) == -1 and b"finite" in lib.kb_last_error()
    assert lib.kb_add_rigid_object(h, -1, T.ctypes.data_as(dp)) == -1
    assert lib.kb_add_terrain(h, -3) == -1 and lib.kb_add_terrain(h, 0) == 0
    par = np.array([-1, 0], dtype=np.int32); lt = np.array([0, 3], dtype=np.uint8)
    ax = np.array([0, 0, 1, 0, 0, 1.0]); T0 = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), 2); q = np.zeros(2)
    args = lambda lt_, q_: (h, 2, par.ctypes.data_as(ip), lt_.ctypes.data_as(_capi.c_uint8_p), ax.ctypes.data_as(dp), T0.ctypes.data_as(dp),
                            q_.ctypes.data_as(dp), q_.ctypes.data_as(dp))
    assert lib.kb_robot_create(*args(lt, q)) == -1 and b"type" in lib.kb_last_error()
    lt[1] = 1; qn = np.array([0.0, np.nan])
    assert lib.kb_robot_create(*args(lt, qn)) == -1 and b"NaN" in lib.kb_last_error()
    assert lib.kb_robot_create(*args(lt, q)) == 0
    lib.kb_engine_destroy(h)
