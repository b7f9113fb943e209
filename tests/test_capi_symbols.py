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
