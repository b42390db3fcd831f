"""The C-ABI library: loads without a GPU, exports every symbol include/dslam_b200.h declares, and refuses to compute
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from direct_stereo_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dslam_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dslam_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/dslam_b200.h is not exported" % n
        assert n in _lib.SIGNATURES, "symbol %s has no ctypes signature in _lib.py" % n
    for n in _lib.SIGNATURES:
        assert n in names, "%s is bound in _lib.py but not declared in the header" % n


def test_exports_match_nm():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dslam_[A-Za-z0-9_]+)", out))
    assert exported == set(declared_symbols())


def test_header_compiles_as_c():
    """include/dslam_b200.h is a plain C header (extern "C" only under __cplusplus)."""
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.dslam_version() >= 100
    assert isinstance(lib.dslam_last_error(), bytes)
    assert lib.dslam_device_count(None) == _lib.EINVAL
    assert b"null" in lib.dslam_last_error()


def test_no_cpu_fallback():
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    p = C.c_void_p()
    rc = lib.dslam_session_create(0, C.byref(p))
    assert rc == _lib.ENODEVICE and not p.value
    from direct_stereo_slam_b200 import api

    with pytest.raises(_lib.DslamError) as e:
        api.Session(0)
    assert e.value.code == _lib.ENODEVICE


def test_library_is_in_tree_and_has_no_link_time_cuda_driver_dependency():
    assert os.path.dirname(_lib.LIB_PATH) == os.path.join(ROOT, "direct_stereo_slam_b200")
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libnccl" not in out and "libcudart" not in out


def test_argument_validation_needs_no_device():
    """Bad arguments are rejected before anything touches CUDA (same error behaviour with and without a GPU)."""
    lib = _lib.load()
    assert lib.dslam_sc_set_scan_kernel(5) == _lib.EINVAL and b"flavour" in lib.dslam_last_error()
    assert lib.dslam_sc_set_scan_kernel(-1) == _lib.EINVAL
    for f in (1, 2, 3, 4, 0):  # process-wide knob, no device needed; leave it on "auto"
        assert lib.dslam_sc_set_scan_kernel(f) == 0
    assert lib.dslam_frame_upload_batch(0, None, None) == _lib.EINVAL
    assert lib.dslam_frame_upload_batch(2, None, None) == _lib.EINVAL
    assert lib.dslam_frame_build_batch(0, None, None, 4) == _lib.EINVAL
    assert lib.dslam_frame_upload(None, None) == _lib.EINVAL
