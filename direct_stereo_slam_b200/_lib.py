"""ctypes binding of libdslam_b200.so (the C ABI declared in include/dslam_b200.h).

The library is the product: if it is missing this module raises — there is no Python or CPU fallback.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSLAM_LIB=<path> loads another build of the same library (A/B runs of kernel variants, tools/build_variants.sh)
LIB_PATH = os.environ.get("DSLAM_LIB") or os.path.join(_HERE, "libdslam_b200.so")
CSRC = os.path.join(_HERE, "csrc")

c_f = C.POINTER(C.c_float)
c_d = C.POINTER(C.c_double)
c_i = C.POINTER(C.c_int)
c_u64 = C.POINTER(C.c_ulonglong)
c_pp = C.POINTER(C.c_void_p)
c_fpp = C.POINTER(c_f)
vp = C.c_void_p

OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, ESTATE, ENCCL, ETIMEOUT = 0, -1, -2, -3, -4, -5, -6, -7
_CODES = {EINVAL: "DSLAM_EINVAL", ENODEVICE: "DSLAM_ENODEVICE", ECUDA: "DSLAM_ECUDA", ENOMEM: "DSLAM_ENOMEM", ESTATE: "DSLAM_ESTATE",
          ENCCL: "DSLAM_ENCCL", ETIMEOUT: "DSLAM_ETIMEOUT"}


class DslamError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("%s (%d): %s" % (_CODES.get(code, "DSLAM_E?"), code, text))
        self.code = code


# every entry point of include/dslam_b200.h: name -> argtypes (all return int unless noted)
SIGNATURES = {
    "dslam_version": [],
    "dslam_last_error": [],
    "dslam_device_count": [c_i],
    "dslam_session_create": [C.c_int, c_pp],
    "dslam_session_destroy": [vp],
    "dslam_session_sync": [vp],
    "dslam_session_stream": [vp, c_pp],
    "dslam_session_launch_count": [vp, C.POINTER(C.c_longlong)],
    "dslam_session_mark": [vp, C.c_int],
    "dslam_session_elapsed_ms": [vp, c_f],
    "dslam_session_profile": [vp, C.c_int],
    "dslam_session_profile_read": [vp, c_d],
    "dslam_session_host_times": [vp, c_d],
    "dslam_host_alloc": [C.c_ulonglong, c_pp],
    "dslam_host_free": [vp],
    "dslam_frame_create": [vp, C.c_int, C.c_int, C.c_int, c_pp],
    "dslam_frame_destroy": [vp],
    "dslam_plan_eval_launch": [C.c_int, c_i, C.c_int, C.c_int, c_i, c_i, c_i],
    "dslam_frame_upload": [vp, c_f],
    "dslam_frame_upload_batch": [C.c_int, c_pp, c_pp],
    "dslam_frame_build": [vp, c_f],
    "dslam_frame_build_batch": [C.c_int, c_pp, c_f, C.c_int],
    "dslam_frame_download": [vp, c_fpp, c_fpp],
    "dslam_frame_wait_host": [vp],
    "dslam_frame_make_images": [vp, c_f, c_f, c_fpp, c_fpp],
    "dslam_ctx_create": [vp, C.c_int, C.c_int, C.c_int, c_f, c_f, c_d, c_pp],
    "dslam_ctx_destroy": [vp],
    "dslam_ctx_make_K": [vp, c_f],
    "dslam_ctx_set_affine_mode": [vp, C.c_int, C.c_int],
    "dslam_ref_upload": [vp, C.c_int, C.c_int, c_f, c_f, c_f, c_f],
    "dslam_ref_set_affine": [vp, C.c_float, C.c_double, C.c_double],
    "dslam_ref_scale_idepth": [vp, C.c_float],
    "dslam_ref_build": [vp, vp, C.c_int, c_i, c_i, c_f, c_f, c_i],
    "dslam_ref_download": [vp, C.c_int, c_i, c_f, c_f, c_f, c_f],
    "dslam_pose_eval": [vp, vp, C.c_float, C.c_int, C.c_int, c_d, c_d, C.c_float, c_d, c_d, c_d, c_i, c_d],
    "dslam_scale_eval": [vp, vp, C.c_int, C.c_int, c_f, C.c_float, c_f, c_d, c_i, c_d],
    "dslam_track_newest_coarse": [vp, vp, C.c_float, c_d, c_d, C.c_int, c_d, c_d, c_d, c_i],
    "dslam_optimize_scale": [vp, vp, c_f, C.c_int, c_f],
    "dslam_optimize_scale_multi": [vp, vp, C.c_int, c_f, C.c_int, c_f],
    "dslam_track_newest_coarse_multi": [vp, vp, C.c_float, C.c_int, c_d, c_d, C.c_int, c_d, c_d, c_d, c_i],
    "dslam_track_newest_coarse_batch": [C.c_int, c_pp, c_pp, c_f, c_d, c_d, C.c_int, c_d, c_d, c_d, c_i],
    "dslam_optimize_scale_batch": [C.c_int, c_pp, c_pp, c_f, C.c_int, c_f],
    "dslam_track_new_coarse": [vp, vp, C.c_float, C.c_int, c_d, c_d, C.c_int, c_d, C.c_double, c_d, c_d, c_d, c_d, c_i, c_i],
    "dslam_lm_batch": [C.c_int, c_pp, c_pp, c_f, c_d, c_d, C.c_int, c_d, c_d, c_d, c_i, C.c_int, c_pp, c_pp, c_f, C.c_int, c_f],
    "dslam_pe_create": [vp, C.c_int, C.c_int, C.c_int, c_pp],
    "dslam_pe_destroy": [vp],
    "dslam_pe_set_affine_mode": [vp, C.c_int, C.c_int],
    "dslam_pe_set_points": [vp, C.c_int, c_d, c_f, C.c_float],
    "dslam_pe_eval": [vp, vp, C.c_float, c_f, C.c_int, c_d, c_d, C.c_float, c_d, c_d, c_d, c_i, c_d],
    "dslam_pe_estimate": [vp, vp, C.c_float, c_f, C.c_int, c_d, c_f, c_i, c_i],
    "dslam_pe_get_trace": [vp, c_d, C.c_int, c_i],
    "dslam_get_trace": [vp, c_d, C.c_int, c_i],
    "dslam_ctx_counters": [vp, C.POINTER(C.c_longlong)],
    "dslam_sc_create": [vp, C.c_int, C.c_int, C.c_int, c_pp],
    "dslam_sc_create_ex": [vp, C.c_int, C.c_int, C.c_int, C.c_int, c_pp],
    "dslam_sc_destroy": [vp],
    "dslam_sc_add64": [vp, C.c_int, c_f, c_d, c_i],
    "dslam_sc_save": [vp, C.c_char_p],
    "dslam_sc_load": [vp, C.c_char_p, c_i],
    "dslam_sc_search_sc64": [vp, C.c_int, c_d, c_i, C.c_int, c_i, c_f],
    "dslam_sc_query64": [vp, C.c_int, c_f, c_d, C.c_float, C.c_int, c_i, c_f],
    "dslam_sc_exchange_mode": [vp, c_i],
    "dslam_sc_add": [vp, C.c_int, c_f, c_f, c_i],
    "dslam_sc_add_sparse": [vp, c_f, c_i, c_d, C.c_int, C.c_int],
    "dslam_sc_generate": [vp, c_d, C.c_int, C.c_double, c_f, c_f, c_d, c_d, C.c_int, C.c_int],
    "dslam_sc_size": [vp, c_i],
    "dslam_sc_search_ringkey": [vp, C.c_int, c_f, C.c_int, C.c_float, C.c_int, c_i, c_f],
    "dslam_sc_search_sc": [vp, C.c_int, c_f, c_i, C.c_int, c_i, c_f],
    "dslam_sc_query": [vp, C.c_int, c_f, c_f, C.c_float, C.c_int, c_i, c_f],
    "dslam_sc_query_keys": [vp, C.c_int, c_f, c_f, C.c_float, C.c_int, c_u64],
    "dslam_sc_decode_key": [C.c_ulonglong, c_i, c_f],
    "dslam_sc_unique_id": [C.POINTER(C.c_ubyte)],
    "dslam_sc_comm_init": [vp, C.POINTER(C.c_ubyte), C.c_int, C.c_int],
    "dslam_sc_last_scan_ms": [vp, c_f],
    "dslam_sc_set_scan_kernel": [C.c_int],
}

_lib = None


def build(verbose=False):
    """Compile libdslam_b200.so for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-C", CSRC, "-j8"], check=True, stdout=out)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libdslam_b200.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C direct_stereo_slam_b200/csrc`; there is no fallback path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_char_p if name == "dslam_last_error" else C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise DslamError(rc, load().dslam_last_error().decode("utf-8", "replace"))
    return rc


def device_count():
    n = C.c_int(0)
    rc = load().dslam_device_count(C.byref(n))
    return n.value if rc == OK else 0
