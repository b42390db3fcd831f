"""ctypes front-end of the CPU oracle (oracle/dslam_oracle.cpp, oracle/sc_generate.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (direct_stereo_slam_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

c_f = C.POINTER(C.c_float)
c_d = C.POINTER(C.c_double)
c_i = C.POINTER(C.c_int)


def _native_dir():
    """-march=native objects are machine specific and must never travel: they live under the temp dir, keyed by the
    oracle sources and the CPU flags of this machine."""
    import hashlib
    import tempfile

    h = hashlib.sha1()
    for f in ("dslam_oracle.cpp", "sc_generate.cpp", "Makefile"):
        h.update(open(os.path.join(_HERE, f), "rb").read())
    try:
        flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0]
    except Exception:
        flags = ""
    h.update(flags.encode())
    return os.path.join(tempfile.gettempdir(), "dslam_oracle_native_" + h.hexdigest()[:16])


def build(native=False, quiet=True):
    """Compile the oracle with the committed Makefile; returns the path of the shared library."""
    out = subprocess.DEVNULL if quiet else None
    if native:
        d = _native_dir()
        subprocess.run(["make", "-C", _HERE, "native", "B=" + d], check=True, stdout=out)
        return os.path.join(d, "libdslam_oracle_native.so")
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=out)
    return os.path.join(_BUILD, "libdslam_oracle.so")


def _fp(a):
    return a.ctypes.data_as(c_f)


def _dp(a):
    return a.ctypes.data_as(c_d)


def _ip(a):
    return a.ctypes.data_as(c_i)


def level_sizes(w, h, levels):
    return [(w >> l, h >> l) for l in range(levels)]


def level_offsets(w, h, levels):
    off = [0]
    for wl, hl in level_sizes(w, h, levels):
        off.append(off[-1] + wl * hl)
    return off


def pyr_levels_used(w, h, max_levels=6):
    """setGlobalCalib's rule, deps:dso/src/util/globalCalib.cpp:45-56."""
    lv = 1
    while w % 2 == 0 and h % 2 == 0 and w * h > 5000 and lv < max_levels:
        w //= 2
        h //= 2
        lv += 1
    return lv


class Oracle:
    def __init__(self, native=False, path=None):
        if path is None:
            path = build(native=native)  # make decides whether anything is out of date
        self.path = path
        L = self.lib = C.CDLL(path)
        L.orc_make_images.argtypes = [c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f]
        L.orc_tracker_create.restype = C.c_void_p
        L.orc_tracker_create.argtypes = [C.c_int, C.c_int, C.c_int, c_f, c_f, c_d]
        L.orc_tracker_destroy.argtypes = [C.c_void_p]
        L.orc_tracker_get_K.argtypes = [C.c_void_p, C.c_int, c_f]
        L.orc_tracker_set_aff_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_tracker_set_res_acc_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_tracker_set_ref_level.argtypes = [C.c_void_p, C.c_int, C.c_int, c_f, c_f, c_f, c_f]
        L.orc_tracker_get_ref_level.restype = C.c_int
        L.orc_tracker_get_ref_level.argtypes = [C.c_void_p, C.c_int, c_f, c_f, c_f, c_f]
        L.orc_tracker_set_ref_aff.argtypes = [C.c_void_p, C.c_float, C.c_double, C.c_double]
        L.orc_tracker_scale_idepth.argtypes = [C.c_void_p, C.c_float]
        L.orc_tracker_set_new_frame.argtypes = [C.c_void_p, c_f, C.c_float]
        L.orc_tracker_set_right_frame.argtypes = [C.c_void_p, c_f]
        L.orc_tracker_make_coarse_depth.argtypes = [C.c_void_p, C.c_int, c_i, c_i, c_f, c_f, c_f]
        L.orc_se3_exp.argtypes = [c_d, c_d]
        L.orc_se3_mul.argtypes = [c_d, c_d, c_d]
        L.orc_se3_R.argtypes = [c_d, c_d]
        L.orc_ldlt_solve.argtypes = [C.c_int, c_d, c_d, c_d]
        L.orc_calc_res_pose.restype = C.c_int
        L.orc_calc_res_pose.argtypes = [C.c_void_p, C.c_int, c_d, C.c_double, C.c_double, C.c_float, c_d]
        L.orc_get_warped.restype = C.c_int
        L.orc_get_warped.argtypes = [C.c_void_p, C.c_int, c_f]
        L.orc_calc_gs_pose.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, c_d, c_d, c_d]
        L.orc_pose_rows.restype = C.c_int
        L.orc_pose_rows.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, c_f, c_f]
        L.orc_scale_rows.restype = C.c_int
        L.orc_scale_rows.argtypes = [C.c_void_p, C.c_int, C.c_float, c_f, c_f, c_f]
        L.orc_accumulator9.argtypes = [c_f, c_f, C.c_int, c_f]
        L.orc_scale_accumulator.argtypes = [c_f, c_f, c_f, C.c_int, c_f]
        L.orc_track_newest_coarse.restype = C.c_int
        L.orc_track_newest_coarse.argtypes = [C.c_void_p, C.c_int, c_d, c_d, C.c_int, c_d, c_d, c_d]
        L.orc_track_new_coarse.restype = C.c_int
        L.orc_track_new_coarse.argtypes = [C.c_void_p, C.c_int, C.c_int, c_d, c_d, C.c_int, c_d, C.c_double, c_d, c_d, c_d, c_d, c_i]
        L.orc_calc_res_scale.restype = C.c_int
        L.orc_calc_res_scale.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, c_d]
        L.orc_calc_gs_scale.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, c_f, c_f, c_d]
        L.orc_optimize_scale.restype = C.c_float
        L.orc_optimize_scale.argtypes = [C.c_void_p, C.c_int, c_f, C.c_int]
        L.orc_get_trace.restype = C.c_int
        L.orc_get_trace.argtypes = [C.c_void_p, c_d, C.c_int]
        L.orc_get_counters.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.orc_pe_create.restype = C.c_void_p
        L.orc_pe_create.argtypes = [C.c_int, C.c_int, C.c_int, c_f]
        L.orc_pe_destroy.argtypes = [C.c_void_p]
        L.orc_pe_set_points.argtypes = [C.c_void_p, C.c_int, c_d, c_f, C.c_float]
        L.orc_pe_set_new_frame.argtypes = [C.c_void_p, c_f, C.c_float]
        L.orc_pe_set_aff_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_pe_calc_res.restype = C.c_int
        L.orc_pe_calc_res.argtypes = [C.c_void_p, C.c_int, C.c_int, c_d, C.c_double, C.c_double, C.c_float, c_d, c_d, c_d, c_d]
        L.orc_pe_estimate.restype = C.c_int
        L.orc_pe_estimate.argtypes = [C.c_void_p, C.c_int, c_d, C.c_int, c_f, c_i]
        L.orc_pe_get_trace.restype = C.c_int
        L.orc_pe_get_trace.argtypes = [C.c_void_p, c_d, C.c_int]
        L.orc_search_sc.argtypes = [c_i, c_d, C.c_int, c_i, c_i, c_d, c_i, C.c_int, C.c_int, c_i, c_f]
        L.orc_search_sc_dense.argtypes = [c_f, c_f, C.c_int, c_i, C.c_int, C.c_int, c_i, c_f]
        L.orc_search_ringkey.restype = C.c_int
        L.orc_search_ringkey.argtypes = [c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_float, c_i, c_f]
        L.orc_sc_generate.restype = C.c_int
        L.orc_sc_generate.argtypes = [c_d, C.c_int, C.c_double, C.c_int, C.c_int, c_f, c_i, c_d, c_d]

    # ---- pyramid ------------------------------------------------------------------------------------
    def make_images(self, img, levels, B256=None):
        """img: (h, w) float32.  Returns (dIp_all [sumP,3] float32, absg_all [sumP] float32)."""
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        tot = level_offsets(w, h, levels)[-1]
        dIp = np.empty((tot, 3), np.float32)
        ag = np.empty(tot, np.float32)
        Bp = None
        if B256 is not None:
            B256 = np.ascontiguousarray(B256, np.float32)
            Bp = _fp(B256)
        self.lib.orc_make_images(_fp(img), w, h, levels, Bp, _fp(dIp), _fp(ag))
        return dIp, ag

    def tracker(self, w, h, levels, K0, K1, T_stereo):
        return OracleTracker(self, w, h, levels, K0, K1, T_stereo)

    def pose_estimator(self, w, h, levels, cam):
        return OraclePoseEstimator(self, w, h, levels, cam)

    # ---- SE3 helpers ----------------------------------------------------------------------------------
    def se3_exp(self, a6):
        a6 = np.ascontiguousarray(a6, np.float64)
        out = np.empty(7, np.float64)
        self.lib.orc_se3_exp(_dp(a6), _dp(out))
        return out

    def se3_mul(self, a7, b7):
        a7 = np.ascontiguousarray(a7, np.float64)
        b7 = np.ascontiguousarray(b7, np.float64)
        out = np.empty(7, np.float64)
        self.lib.orc_se3_mul(_dp(a7), _dp(b7), _dp(out))
        return out

    def se3_R(self, a7):
        a7 = np.ascontiguousarray(a7, np.float64)
        out = np.empty(9, np.float64)
        self.lib.orc_se3_R(_dp(a7), _dp(out))
        return out.reshape(3, 3)

    def ldlt_solve(self, A, rhs):
        n = len(rhs)
        A8 = np.zeros((8, 8), np.float64)
        A8[:n, :n] = A
        rhs = np.ascontiguousarray(rhs, np.float64)
        x = np.zeros(8, np.float64)
        self.lib.orc_ldlt_solve(n, _dp(A8), _dp(rhs), _dp(x))
        return x[:n]

    def accumulator9(self, J, w):
        J = np.ascontiguousarray(J, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        out = np.zeros(45, np.float32)
        self.lib.orc_accumulator9(_fp(J), _fp(w), J.shape[1], _fp(out))
        return out

    def scale_accumulator(self, J, r, w):
        J, r, w = (np.ascontiguousarray(a, np.float32) for a in (J, r, w))
        out = np.zeros(3, np.float32)
        self.lib.orc_scale_accumulator(_fp(J), _fp(r), _fp(w), len(J), _fp(out))
        return out

    # ---- Scan Context ---------------------------------------------------------------------------------
    def sc_generate(self, pts, lidar_range=40.0, num_s=60, num_r=20):
        pts = np.ascontiguousarray(pts, np.float64)
        ringkey = np.empty(num_r, np.float32)
        idx = np.empty(num_s * num_r, np.int32)
        val = np.empty(num_s * num_r, np.float64)
        tfm = np.empty(16, np.float64)
        nnz = self.lib.orc_sc_generate(_dp(pts), len(pts), lidar_range, num_s, num_r, _fp(ringkey), _ip(idx), _dp(val), _dp(tfm))
        return ringkey, idx[:nnz].copy(), val[:nnz].copy(), tfm.reshape(4, 4)

    def search_sc(self, q_idx, q_val, sig_ptr, sig_idx, sig_val, candidates, sc_width=60):
        q_idx = np.ascontiguousarray(q_idx, np.int32)
        q_val = np.ascontiguousarray(q_val, np.float64)
        sig_ptr = np.ascontiguousarray(sig_ptr, np.int32)
        sig_idx = np.ascontiguousarray(sig_idx, np.int32)
        sig_val = np.ascontiguousarray(sig_val, np.float64)
        candidates = np.ascontiguousarray(candidates, np.int32)
        ri = C.c_int(-1)
        rd = C.c_float(0)
        self.lib.orc_search_sc(_ip(q_idx), _dp(q_val), len(q_idx), _ip(sig_ptr), _ip(sig_idx), _dp(sig_val), _ip(candidates),
                               len(candidates), sc_width, C.byref(ri), C.byref(rd))
        return ri.value, rd.value

    def search_sc_dense(self, q, db, candidates=None, sc_width=60):
        q = np.ascontiguousarray(q, np.float32)
        db = np.ascontiguousarray(db, np.float32)
        n_cells = q.shape[-1]
        ri = C.c_int(-1)
        rd = C.c_float(0)
        if candidates is None:
            self.lib.orc_search_sc_dense(_fp(q), _fp(db), n_cells, None, db.shape[0], sc_width, C.byref(ri), C.byref(rd))
        else:
            candidates = np.ascontiguousarray(candidates, np.int32)
            self.lib.orc_search_sc_dense(_fp(q), _fp(db), n_cells, _ip(candidates), len(candidates), sc_width, C.byref(ri), C.byref(rd))
        return ri.value, rd.value

    def search_ringkey(self, q, keys, k=3, thres=0.1):
        q = np.ascontiguousarray(q, np.float32)
        keys = np.ascontiguousarray(keys, np.float32)
        cand = np.empty(k, np.int32)
        dist = np.empty(k, np.float32)
        n = self.lib.orc_search_ringkey(_fp(q), _fp(keys), keys.shape[0], keys.shape[1], k, thres, _ip(cand), _fp(dist))
        return cand[:n].copy(), dist[:n].copy()


class OracleTracker:
    """One dso::TrackerAndScaler restated (src/scale_optimization/TrackerAndScaler.{h,cpp})."""

    SSE, FP64 = 0, 1  # accumulation mode of calcGSSSE*

    def __init__(self, orc, w, h, levels, K0, K1, T_stereo):
        self.o = orc
        self.L = orc.lib
        self.w, self.h, self.levels = w, h, levels
        K0 = np.ascontiguousarray(K0, np.float32)
        K1 = np.ascontiguousarray(K1, np.float32)
        T = np.ascontiguousarray(T_stereo, np.float64).reshape(16)
        self.p = C.c_void_p(self.L.orc_tracker_create(w, h, levels, _fp(K0), _fp(K1), _dp(T)))
        self._keep = {}

    def __del__(self):
        try:
            self.L.orc_tracker_destroy(self.p)
        except Exception:
            pass

    def get_K(self, lvl):
        out = np.empty(17, np.float32)
        self.L.orc_tracker_get_K(self.p, lvl, _fp(out))
        return dict(fx=out[0], fy=out[1], cx=out[2], cy=out[3], Ki=out[4:13].reshape(3, 3).copy(), fx1=out[13], fy1=out[14], cx1=out[15], cy1=out[16])

    def set_res_acc_mode(self, mode):
        """0 = fp32 sequential E / flow sums (reference), 1 = fp64 sums of the same fp32 terms (what the CUDA path does)."""
        self.L.orc_tracker_set_res_acc_mode(self.p, mode)

    def set_aff_mode(self, a, b):
        self.L.orc_tracker_set_aff_mode(self.p, a, b)

    def set_ref_level(self, lvl, u, v, idepth, color):
        u, v, idepth, color = (np.ascontiguousarray(a, np.float32) for a in (u, v, idepth, color))
        self.L.orc_tracker_set_ref_level(self.p, lvl, len(u), _fp(u), _fp(v), _fp(idepth), _fp(color))

    def get_ref_level(self, lvl):
        n = self.L.orc_tracker_get_ref_level(self.p, lvl, None, None, None, None)
        arrs = [np.empty(n, np.float32) for _ in range(4)]
        self.L.orc_tracker_get_ref_level(self.p, lvl, *[_fp(a) for a in arrs])
        return arrs

    def set_ref_aff(self, exposure, a, b):
        self.L.orc_tracker_set_ref_aff(self.p, exposure, a, b)

    def scale_idepth(self, s):
        self.L.orc_tracker_scale_idepth(self.p, s)

    def set_new_frame(self, dIp_all, exposure=1.0):
        self._keep["new"] = dIp_all
        self.L.orc_tracker_set_new_frame(self.p, _fp(dIp_all), exposure)

    def set_right_frame(self, dIp_all):
        self._keep["right"] = dIp_all
        self.L.orc_tracker_set_right_frame(self.p, _fp(dIp_all))

    def make_coarse_depth(self, pu, pv, pid, pweight, dIp_ref_all):
        pu = np.ascontiguousarray(pu, np.int32)
        pv = np.ascontiguousarray(pv, np.int32)
        pid = np.ascontiguousarray(pid, np.float32)
        pweight = np.ascontiguousarray(pweight, np.float32)
        self.L.orc_tracker_make_coarse_depth(self.p, len(pu), _ip(pu), _ip(pv), _fp(pid), _fp(pweight), _fp(dIp_ref_all))

    def calc_res_pose(self, lvl, pose7, aff, cutoff=20.0):
        pose7 = np.ascontiguousarray(pose7, np.float64)
        res = np.empty(6, np.float64)
        n = self.L.orc_calc_res_pose(self.p, lvl, _dp(pose7), aff[0], aff[1], cutoff, _dp(res))
        return res, n

    def get_warped(self, which=0):
        n = self.L.orc_get_warped(self.p, which, None)
        out = np.empty((8, n), np.float32)
        self.L.orc_get_warped(self.p, which, _fp(out))
        return out

    def calc_gs_pose(self, lvl, mode, aff):
        H = np.empty(64, np.float64)
        b = np.empty(8, np.float64)
        acc = np.empty(45, np.float64)
        self.L.orc_calc_gs_pose(self.p, lvl, mode, aff[0], aff[1], _dp(H), _dp(b), _dp(acc))
        return H.reshape(8, 8), b, acc

    def pose_rows(self, lvl, aff):
        """(J [9, n], w [n]) of the last calc_res_pose, as calcGSSSEPose forms them."""
        n = self.L.orc_pose_rows(self.p, lvl, aff[0], aff[1], None, None)
        J = np.zeros((9, n), np.float32)
        w = np.zeros(n, np.float32)
        self.L.orc_pose_rows(self.p, lvl, aff[0], aff[1], _fp(J), _fp(w))
        return J, w

    def scale_rows(self, lvl, scale):
        n = self.L.orc_scale_rows(self.p, lvl, scale, None, None, None)
        J, r, w = (np.zeros(n, np.float32) for _ in range(3))
        self.L.orc_scale_rows(self.p, lvl, scale, _fp(J), _fp(r), _fp(w))
        return J, r, w

    def track_newest_coarse(self, mode, pose7, aff, coarsest, min_res=None):
        pose7 = np.array(pose7, np.float64)
        aff = np.array(aff, np.float64)
        min_res = np.full(5, np.nan) if min_res is None else np.ascontiguousarray(min_res, np.float64)
        last = np.empty(5, np.float64)
        flow = np.empty(3, np.float64)
        ok = self.L.orc_track_newest_coarse(self.p, mode, _dp(pose7), _dp(aff), coarsest, _dp(min_res), _dp(last), _dp(flow))
        return bool(ok), pose7, aff, last, flow

    def track_new_coarse(self, mode, tries7, aff_init, coarsest, last_coarse_rmse, re_track_threshold=1.5):
        """The retry loop of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:192-252), sequential like the reference."""
        tries7 = np.ascontiguousarray(np.atleast_2d(tries7), np.float64)
        aff0 = np.ascontiguousarray(aff_init, np.float64)
        last = np.ascontiguousarray(last_coarse_rmse, np.float64)
        pose, aff, ach, flow = np.empty(7), np.empty(2), np.empty(5), np.empty(3)
        ntry = C.c_int(0)
        good = self.L.orc_track_new_coarse(self.p, mode, len(tries7), _dp(tries7), _dp(aff0), coarsest, _dp(last), re_track_threshold, _dp(pose),
                                           _dp(aff), _dp(ach), _dp(flow), C.byref(ntry))
        return dict(pose=pose, aff=aff, achievedRes=ach, flow=flow, haveOneGood=bool(good), tryIterations=ntry.value)

    def calc_res_scale(self, lvl, scale, cutoff=20.0):
        res = np.empty(6, np.float64)
        n = self.L.orc_calc_res_scale(self.p, lvl, scale, cutoff, _dp(res))
        return res, n

    def calc_gs_scale(self, lvl, mode, scale):
        H = C.c_float(0)
        b = C.c_float(0)
        acc = np.empty(3, np.float64)
        self.L.orc_calc_gs_scale(self.p, lvl, mode, scale, C.byref(H), C.byref(b), _dp(acc))
        return H.value, b.value, acc

    def optimize_scale(self, mode, scale, coarsest):
        s = C.c_float(scale)
        rmse = self.L.orc_optimize_scale(self.p, mode, C.byref(s), coarsest)
        return rmse, s.value

    def trace(self):
        n = self.L.orc_get_trace(self.p, None, 0)
        out = np.zeros((n, 15), np.float64)
        if n:
            self.L.orc_get_trace(self.p, _dp(out), n)
        return out

    def counters(self):
        out = (C.c_long * 2)()
        self.L.orc_get_counters(self.p, out)
        return out[0], out[1]


class ReferencePieces:
    """oracle/_ref/libdslam_ref.so: the reference's OWN Accumulator9 / ScaleAccumulator / search_place.h compiled in place by
    oracle/ref_build.py.  Used to pin the restatements above."""

    def __init__(self, path=None):
        path = path or os.path.join(_HERE, "_ref", "libdslam_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_accumulator9.argtypes = [c_f, c_f, C.c_int, c_f]
        L.ref_scale_accumulator.argtypes = [c_f, c_f, c_f, C.c_int, c_f]
        L.ref_search_sc.argtypes = [c_i, c_d, C.c_int, c_i, c_i, c_d, C.c_int, c_i, C.c_int, C.c_int, c_i, c_f]
        L.ref_search_ringkey_sequence.argtypes = [c_f, C.c_int, C.c_int, c_i, c_i]

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libdslam_ref.so"))

    def accumulator9(self, J, w):
        J = np.ascontiguousarray(J, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        out = np.zeros(45, np.float32)
        self.lib.ref_accumulator9(_fp(J), _fp(w), J.shape[1], _fp(out))
        return out

    def scale_accumulator(self, J, r, w):
        J, r, w = (np.ascontiguousarray(a, np.float32) for a in (J, r, w))
        out = np.zeros(3, np.float32)
        self.lib.ref_scale_accumulator(_fp(J), _fp(r), _fp(w), len(J), _fp(out))
        return out

    def search_sc(self, q_idx, q_val, sig_ptr, sig_idx, sig_val, candidates, sc_width=60):
        q_idx = np.ascontiguousarray(q_idx, np.int32)
        q_val = np.ascontiguousarray(q_val, np.float64)
        sig_ptr = np.ascontiguousarray(sig_ptr, np.int32)
        sig_idx = np.ascontiguousarray(sig_idx, np.int32)
        sig_val = np.ascontiguousarray(sig_val, np.float64)
        candidates = np.ascontiguousarray(candidates, np.int32)
        ri = C.c_int(-1)
        rd = C.c_float(0)
        self.lib.ref_search_sc(_ip(q_idx), _dp(q_val), len(q_idx), _ip(sig_ptr), _ip(sig_idx), _dp(sig_val), len(sig_ptr) - 1, _ip(candidates),
                               len(candidates), sc_width, C.byref(ri), C.byref(rd))
        return ri.value, rd.value

    def search_ringkey_sequence(self, keys):
        """Run the keys through the reference's search_ringkey in order (once per process: it keeps static state)."""
        keys = np.ascontiguousarray(keys, np.float32)
        n, dim = keys.shape
        cand = np.full((n, 3), -1, np.int32)
        ncand = np.zeros(n, np.int32)
        self.lib.ref_search_ringkey_sequence(_fp(keys), n, dim, _ip(cand), _ip(ncand))
        return cand, ncand


class ReferenceTracker:
    """oracle/_ref/libdslam_ref_tracker.so: the reference's OWN TrackerAndScaler.cpp hot path (constructor, makeK,
    makeCoarseDepthL0, trackNewestCoarse, calcResPose, calcGSSSEPose, optimizeScale, calcResScale, calcGSSSEScale) compiled
    in place against the Eigen / Sophus / DSO stand-ins of oracle/shim (oracle/ref_build.py)."""

    OPT_PATH = os.path.join(_HERE, "_ref", "libdslam_ref_tracker_opt.so")  # -O3 build for timing

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libdslam_ref_tracker.so"))

    def __init__(self, w, h, levels, K0, K1, T_stereo, path=None):
        path = path or os.path.join(_HERE, "_ref", "libdslam_ref_tracker.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.L = C.CDLL(path)
        L.reft_create.restype = C.c_void_p
        L.reft_create.argtypes = [C.c_int, C.c_int, C.c_int, c_f, c_f, c_d]
        L.reft_destroy.argtypes = [C.c_void_p]
        L.reft_set_aff_mode.argtypes = [C.c_float, C.c_float]
        L.reft_set_ref.argtypes = [C.c_void_p, c_f, C.c_int, c_i, c_i, c_f, c_f, C.c_float, C.c_double, C.c_double]
        L.reft_get_ref_level.restype = C.c_int
        L.reft_get_ref_level.argtypes = [C.c_void_p, C.c_int, c_f, c_f, c_f, c_f]
        L.reft_scale_idepth.argtypes = [C.c_void_p, C.c_float]
        L.reft_set_new_frame.argtypes = [C.c_void_p, c_f, C.c_float]
        L.reft_set_right_frame.argtypes = [C.c_void_p, c_f]
        L.reft_calc_res_pose.restype = C.c_int
        L.reft_calc_res_pose.argtypes = [C.c_void_p, C.c_int, c_d, C.c_double, C.c_double, C.c_float, c_d]
        L.reft_calc_gs_pose.argtypes = [C.c_void_p, C.c_int, c_d, C.c_double, C.c_double, c_d, c_d]
        L.reft_get_warped.restype = C.c_int
        L.reft_get_warped.argtypes = [C.c_void_p, C.c_int, c_f]
        L.reft_track.restype = C.c_int
        L.reft_track.argtypes = [C.c_void_p, c_d, c_d, C.c_int, c_d, c_d, c_d]
        L.reft_calc_res_scale.restype = C.c_int
        L.reft_calc_res_scale.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, c_d]
        L.reft_calc_gs_scale.argtypes = [C.c_void_p, C.c_int, C.c_float, c_f, c_f]
        L.reft_optimize_scale.restype = C.c_float
        L.reft_optimize_scale.argtypes = [C.c_void_p, c_f, C.c_int]
        L.reft_get_K.argtypes = [C.c_void_p, C.c_int, c_f]
        L.refimg_make_images.argtypes = [c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f]
        K0 = np.ascontiguousarray(K0, np.float32)
        K1 = np.ascontiguousarray(K1, np.float32)
        T = np.ascontiguousarray(T_stereo, np.float64).reshape(16)
        self.levels = levels
        self.p = C.c_void_p(L.reft_create(w, h, levels, _fp(K0), _fp(K1), _dp(T)))
        self._keep = {}

    def __del__(self):
        try:
            self.L.reft_destroy(self.p)
        except Exception:
            pass

    def set_aff_mode(self, a, b):
        self.L.reft_set_aff_mode(a, b)

    @staticmethod
    def weight_from_hdif(hdif):
        """weight = sqrtf(1e-3 / (HdiF + 1e-12))  (TrackerAndScaler.cpp:160) in the C types of that expression."""
        hdif = np.asarray(hdif, np.float32)
        return np.sqrt((1e-3 / (hdif.astype(np.float64) + 1e-12)).astype(np.float32)).astype(np.float32)

    def set_ref(self, dIp_ref_all, pu, pv, pid, hdif, exposure=1.0, a=0.0, b=0.0):
        pu = np.ascontiguousarray(pu, np.int32)
        pv = np.ascontiguousarray(pv, np.int32)
        pid = np.ascontiguousarray(pid, np.float32)
        hdif = np.ascontiguousarray(hdif, np.float32)
        self._keep["ref"] = (dIp_ref_all, pu, pv, pid, hdif)
        self.L.reft_set_ref(self.p, _fp(dIp_ref_all), len(pu), _ip(pu), _ip(pv), _fp(pid), _fp(hdif), exposure, a, b)

    def get_ref_level(self, lvl):
        n = self.L.reft_get_ref_level(self.p, lvl, None, None, None, None)
        arrs = [np.empty(n, np.float32) for _ in range(4)]
        self.L.reft_get_ref_level(self.p, lvl, *[_fp(x) for x in arrs])
        return arrs

    def scale_idepth(self, s):
        self.L.reft_scale_idepth(self.p, s)

    def set_new_frame(self, dIp_all, exposure=1.0):
        self._keep["new"] = dIp_all
        self.L.reft_set_new_frame(self.p, _fp(dIp_all), exposure)

    def set_right_frame(self, dIp_all):
        self._keep["right"] = dIp_all
        self.L.reft_set_right_frame(self.p, _fp(dIp_all))

    def calc_res_pose(self, lvl, pose7, aff, cutoff=20.0):
        pose7 = np.ascontiguousarray(pose7, np.float64)
        res = np.empty(6, np.float64)
        n = self.L.reft_calc_res_pose(self.p, lvl, _dp(pose7), aff[0], aff[1], cutoff, _dp(res))
        return res, n

    def calc_gs_pose(self, lvl, pose7, aff):
        pose7 = np.ascontiguousarray(pose7, np.float64)
        H = np.empty(64, np.float64)
        b = np.empty(8, np.float64)
        self.L.reft_calc_gs_pose(self.p, lvl, _dp(pose7), aff[0], aff[1], _dp(H), _dp(b))
        return H.reshape(8, 8), b

    def get_warped(self, which=0):
        n = self.L.reft_get_warped(self.p, which, None)
        out = np.empty((8, n), np.float32)
        self.L.reft_get_warped(self.p, which, _fp(out))
        return out

    def track_newest_coarse(self, pose7, aff, coarsest, min_res=None):
        pose7 = np.array(pose7, np.float64)
        aff = np.array(aff, np.float64)
        min_res = np.full(5, np.nan) if min_res is None else np.ascontiguousarray(min_res, np.float64)
        last = np.empty(5, np.float64)
        flow = np.empty(3, np.float64)
        ok = self.L.reft_track(self.p, _dp(pose7), _dp(aff), coarsest, _dp(min_res), _dp(last), _dp(flow))
        return bool(ok), pose7, aff, last, flow

    def calc_res_scale(self, lvl, scale, cutoff=20.0):
        res = np.empty(6, np.float64)
        n = self.L.reft_calc_res_scale(self.p, lvl, scale, cutoff, _dp(res))
        return res, n

    def calc_gs_scale(self, lvl, scale):
        H = C.c_float(0)
        b = C.c_float(0)
        self.L.reft_calc_gs_scale(self.p, lvl, scale, C.byref(H), C.byref(b))
        return H.value, b.value

    def optimize_scale(self, scale, coarsest):
        s = C.c_float(scale)
        rmse = self.L.reft_optimize_scale(self.p, C.byref(s), coarsest)
        return rmse, s.value

    def get_K(self, lvl):
        out = np.empty(17, np.float32)
        self.L.reft_get_K(self.p, lvl, _fp(out))
        return out

    def run_frames(self, k0, k1, phase, kf_every, img_new0, img_new1, img_right, pose_init_2x7, coarsest):
        """Frames [k0, k1) of one stereo stream entirely on the C side (reft_run_frames): makeImages(left) + trackNewestCoarse,
        on keyframes makeImages(right) + optimizeScale(1.0).  Returns (#frames tracked OK, last pose7, last scale)."""
        imgs = [np.ascontiguousarray(a, np.float32) for a in (img_new0, img_new1, img_right)]
        init = np.ascontiguousarray(pose_init_2x7, np.float64).reshape(14)
        pose = np.zeros(7, np.float64)
        scale = C.c_float(0)
        self.L.reft_run_frames.restype = C.c_int
        self.L.reft_run_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_d, C.c_int, c_d, C.POINTER(C.c_float), c_d]
        self.last_phase_s = np.zeros(3, np.float64)  # seconds in makeImages(left), trackNewestCoarse, keyframe work (right pyramid + scale)
        good = self.L.reft_run_frames(self.p, k0, k1, phase, kf_every, _fp(imgs[0]), _fp(imgs[1]), _fp(imgs[2]), _dp(init), coarsest, _dp(pose),
                                      C.byref(scale), _dp(self.last_phase_s))
        return good, pose, scale.value


_REF_LIBS = {}


def reference_make_images(img, levels, B256=None, path=None):
    """The reference's own FrameHessian::makeImages (deps:dso HessianBlocks.cpp:128-191 compiled in place); same output layout
    as Oracle.make_images (first / last row of dx, dy, absSquaredGrad zeroed: the reference leaves them uninitialised)."""
    path = path or os.path.join(_HERE, "_ref", "libdslam_ref_tracker.so")
    if path not in _REF_LIBS:
        _REF_LIBS[path] = C.CDLL(path)
    L = _REF_LIBS[path]
    L.refimg_make_images.argtypes = [c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f]
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    tot = level_offsets(w, h, levels)[-1]
    dIp = np.empty((tot, 3), np.float32)
    ag = np.empty(tot, np.float32)
    Bp = None
    if B256 is not None:
        B256 = np.ascontiguousarray(B256, np.float32)
        Bp = _fp(B256)
    L.refimg_make_images(_fp(img), w, h, levels, Bp, _fp(dIp), _fp(ag))
    return dIp, ag


class ReferenceScanContext:
    """oracle/_ref/libdslam_ref_sc.so: the reference's own ScanContext.cpp (align_points_PCA + ScanContext::generate) compiled
    in place against oracle/shim_sc (its SelfAdjointEigenSolver is oracle/jacobi_eig3.h, the solver the oracle uses too)."""

    PATH = os.path.join(_HERE, "_ref", "libdslam_ref_sc.so")

    @staticmethod
    def available():
        return os.path.exists(ReferenceScanContext.PATH)

    def __init__(self):
        self.L = C.CDLL(self.PATH)
        self.L.refsc_generate.restype = C.c_int
        self.L.refsc_generate.argtypes = [c_d, C.c_int, C.c_double, C.c_int, C.c_int, c_f, c_i, c_d, c_d]

    def generate(self, pts, lidar_range=40.0, num_s=60, num_r=20):
        """-> (ringkey [num_r] float32, sig_idx int32 ascending, sig_val float64, tfm_pca_rig 4x4)"""
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        rk = np.empty(num_r, np.float32)
        idx = np.empty(num_s * num_r, np.int32)
        val = np.empty(num_s * num_r, np.float64)
        tfm = np.empty(16, np.float64)
        n = self.L.refsc_generate(_dp(pts), len(pts), lidar_range, num_s, num_r, _fp(rk), _ip(idx), _dp(val), _dp(tfm))
        return rk, idx[:n].copy(), val[:n].copy(), tfm.reshape(4, 4)


class OraclePoseEstimator:
    """dso::PoseEstimator restated (src/loop_closure/pose_estimation/PoseEstimator.cpp)."""

    def __init__(self, orc, w, h, levels, cam):
        self.L = orc.lib
        cam = np.ascontiguousarray(cam, np.float32)
        self.levels = levels
        self.p = C.c_void_p(self.L.orc_pe_create(w, h, levels, _fp(cam)))
        self._keep = {}

    def __del__(self):
        try:
            self.L.orc_pe_destroy(self.p)
        except Exception:
            pass

    def set_points(self, pts, colors, ref_exposure=1.0):
        """pts [n,3] float64; colors [levels, n] float32 (pair.second[lvl])."""
        pts = np.ascontiguousarray(pts, np.float64)
        colors = np.ascontiguousarray(colors, np.float32)
        self.L.orc_pe_set_points(self.p, len(pts), _dp(pts), _fp(colors), ref_exposure)

    def set_new_frame(self, dIp_all, exposure=1.0):
        self._keep["new"] = dIp_all
        self.L.orc_pe_set_new_frame(self.p, _fp(dIp_all), exposure)

    def set_aff_mode(self, a, b):
        self.L.orc_pe_set_aff_mode(self.p, a, b)

    def calc_res(self, lvl, mode, T, aff=(0.0, 0.0), cutoff=20.0):
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        res, H, b, acc = np.empty(6), np.empty(64), np.empty(8), np.empty(45)
        n = self.L.orc_pe_calc_res(self.p, lvl, mode, _dp(T), aff[0], aff[1], cutoff, _dp(res), _dp(H), _dp(b), _dp(acc))
        return res, n, H.reshape(8, 8), b, acc

    def estimate(self, mode, T, coarsest):
        T = np.array(T, np.float64).reshape(16)
        err = C.c_float(0)
        inl = C.c_int(0)
        ok = self.L.orc_pe_estimate(self.p, mode, _dp(T), coarsest, C.byref(err), C.byref(inl))
        return bool(ok), T.reshape(4, 4), err.value, inl.value

    def trace(self):
        n = self.L.orc_pe_get_trace(self.p, None, 0)
        out = np.zeros((n, 15), np.float64)
        if n:
            self.L.orc_pe_get_trace(self.p, _dp(out), n)
        return out


class ReferencePoseEstimator:
    """oracle/_ref/libdslam_ref_pe.so: the reference's own PoseEstimator.cpp compiled in place (oracle/ref_build.py)."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libdslam_ref_pe.so"))

    def __init__(self, w, h, levels, cam):
        L = self.L = C.CDLL(os.path.join(_HERE, "_ref", "libdslam_ref_pe.so"))
        L.refpe_create.restype = C.c_void_p
        L.refpe_create.argtypes = [C.c_int, C.c_int, C.c_int, c_f]
        L.refpe_destroy.argtypes = [C.c_void_p]
        L.refpe_set_aff_mode.argtypes = [C.c_float, C.c_float]
        L.refpe_set_points.argtypes = [C.c_void_p, C.c_int, c_d, c_f, C.c_float]
        L.refpe_set_new_frame.argtypes = [C.c_void_p, c_f, C.c_float]
        L.refpe_estimate.restype = C.c_int
        L.refpe_estimate.argtypes = [C.c_void_p, c_d, C.c_int, c_f]
        L.refpe_calc_res.restype = C.c_int
        L.refpe_calc_res.argtypes = [C.c_void_p, C.c_int, c_d, C.c_double, C.c_double, C.c_float, c_d, c_d, c_d]
        cam = np.ascontiguousarray(cam, np.float32)
        self.p = C.c_void_p(L.refpe_create(w, h, levels, _fp(cam)))
        self._keep = {}

    def __del__(self):
        try:
            self.L.refpe_destroy(self.p)
        except Exception:
            pass

    def set_aff_mode(self, a, b):
        self.L.refpe_set_aff_mode(a, b)

    def set_points(self, pts, colors, ref_exposure=1.0):
        pts = np.ascontiguousarray(pts, np.float64)
        colors = np.ascontiguousarray(colors, np.float32)
        self.L.refpe_set_points(self.p, len(pts), _dp(pts), _fp(colors), ref_exposure)

    def set_new_frame(self, dIp_all, exposure=1.0):
        self._keep["new"] = dIp_all
        self.L.refpe_set_new_frame(self.p, _fp(dIp_all), exposure)

    def calc_res(self, lvl, T, aff=(0.0, 0.0), cutoff=20.0):
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        res, H, b = np.empty(6), np.empty(64), np.empty(8)
        n = self.L.refpe_calc_res(self.p, lvl, _dp(T), aff[0], aff[1], cutoff, _dp(res), _dp(H), _dp(b))
        return res, n, H.reshape(8, 8), b

    def estimate(self, T, coarsest):
        T = np.array(T, np.float64).reshape(16)
        err = C.c_float(0)
        ok = self.L.refpe_estimate(self.p, _dp(T), coarsest, C.byref(err))
        return bool(ok), T.reshape(4, 4), err.value
