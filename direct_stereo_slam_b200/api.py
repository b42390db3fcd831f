"""Host-side mirror of the reference's interface for the hot path, on top of the C ABI.

Names and argument meaning follow the reference (src/scale_optimization/TrackerAndScaler.h:34-137,
deps:dso/src/FullSystem/HessianBlocks.h FrameHessian::makeImages, src/loop_closure/loop_detection/search_place.h):

    FrameHessian.makeImages(color, B256)             -> dIp[l], absSquaredGrad[l]
    TrackerAndScaler(w, h, tfm_vec, K1)              .makeK / .setCoarseTrackingRef / .scaleCoarseDepthL0
    TrackerAndScaler.trackNewestCoarse(...)          -> bool, pose, affine, lastResiduals
    TrackerAndScaler.optimizeScale(fh1, scale, lvl)  -> rmse, scale
    ScanContextDB.search_ringkey / .search_sc / .query

Poses are 7 doubles (qx, qy, qz, qw, tx, ty, tz) == Sophus::SE3d::data(); AffLight is (a, b).
Everything computes on the GPU through libdslam_b200.so; nothing here falls back to numpy.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_d, c_f, c_i, c_u64, check

MAX_LEVELS = 6


def _fp(a):
    return None if a is None else a.ctypes.data_as(c_f)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_d)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_i)


def pyr_levels_used(w, h, max_levels=MAX_LEVELS):
    """setGlobalCalib's level rule (deps:dso/src/util/globalCalib.cpp:45-56)."""
    lv = 1
    while w % 2 == 0 and h % 2 == 0 and w * h > 5000 and lv < max_levels:
        w //= 2
        h //= 2
        lv += 1
    return lv


class Session:
    """One CUDA stream; every object created from it is stream-ordered on it (one host thread at a time)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        p = C.c_void_p()
        check(self.lib.dslam_session_create(device, C.byref(p)))
        self.p = p
        self.device = device
        self._pinned = []

    def close(self):
        if self.p:
            for ptr in self._pinned:
                self.lib.dslam_host_free(ptr)
            self._pinned = []
            self.lib.dslam_session_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.dslam_session_sync(self.p))

    def launch_count(self):
        n = C.c_longlong(0)
        check(self.lib.dslam_session_launch_count(self.p, C.byref(n)))
        return n.value

    def mark(self, which):
        check(self.lib.dslam_session_mark(self.p, which))

    def elapsed_ms(self):
        ms = C.c_float(0)
        check(self.lib.dslam_session_elapsed_ms(self.p, C.byref(ms)))
        return ms.value

    def profile(self, enable):
        check(self.lib.dslam_session_profile(self.p, 1 if enable else 0))

    def profile_read(self):
        """Per-launch CUDA-event timing of the residual kernels since the last read."""
        out = np.zeros(12)
        check(self.lib.dslam_session_profile_read(self.p, _dp(out)))
        return {k: dict(launches=int(out[4 * i]), ms=out[4 * i + 1], points=int(out[4 * i + 2]), max_ms=out[4 * i + 3])
                for i, k in enumerate(("pose", "scale", "mixed"))}

    def host_times(self):
        out = np.zeros(4)
        check(self.lib.dslam_session_host_times(self.p, _dp(out)))
        return dict(prep_ms=out[0], launch_ms=out[1], wait_ms=out[2], launches=int(out[3]))

    def stream(self):
        p = C.c_void_p()
        check(self.lib.dslam_session_stream(self.p, C.byref(p)))
        return p.value

    def pinned(self, shape, dtype=np.float32):
        """numpy array backed by page-locked host memory (cudaHostAlloc) — H2D / D2H of it is truly asynchronous."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        check(self.lib.dslam_host_alloc(max(n, 16), C.byref(ptr)))
        self._pinned.append(ptr)
        buf = (C.c_char * max(n, 16)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class FrameHessian:
    """Device twin of dso::FrameHessian's image pyramid (dIp[], absSquaredGrad[])."""

    def __init__(self, session, w, h, levels=None):
        self.s = session
        self.lib = session.lib
        self.w, self.h = w, h
        self.levels = pyr_levels_used(w, h) if levels is None else levels
        p = C.c_void_p()
        check(self.lib.dslam_frame_create(session.p, w, h, self.levels, C.byref(p)))
        self.p = p
        self.sizes = [(w >> l, h >> l) for l in range(self.levels)]
        self.offsets = [0]
        for wl, hl in self.sizes:
            self.offsets.append(self.offsets[-1] + wl * hl)
        self.dIp_all = None  # [sumP, 3] host copy in the reference layout
        self.absSquaredGrad_all = None
        self.ab_exposure = 1.0
        self._keep = None

    def close(self):
        if self.p:
            self.lib.dslam_frame_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _level_ptrs(self, arr, mul):
        ptrs = (c_f * MAX_LEVELS)()
        base = arr.ctypes.data
        for l in range(self.levels):
            ptrs[l] = C.cast(base + 4 * mul * self.offsets[l], c_f)
        return ptrs

    def alloc_host(self, pinned=True):
        tot = self.offsets[-1]
        if pinned:
            self.dIp_all = self.s.pinned((tot, 3), np.float32)
            self.absSquaredGrad_all = self.s.pinned((tot,), np.float32)
        else:
            self.dIp_all = np.empty((tot, 3), np.float32)
            self.absSquaredGrad_all = np.empty(tot, np.float32)

    def makeImages(self, color, B256=None, host=True, wait=True):
        """FrameHessian::makeImages (deps:dso/src/FullSystem/HessianBlocks.cpp:128-191): upload + build (+ download)."""
        color = np.ascontiguousarray(color, np.float32)
        assert color.size == self.w * self.h
        Bp = None
        if B256 is not None:
            B256 = np.ascontiguousarray(B256, np.float32)
            assert B256.size == 256
            Bp = _fp(B256)
        self._keep = (color, B256)
        if host:
            if self.dIp_all is None:
                self.alloc_host()
            d = self._level_ptrs(self.dIp_all, 3)
            a = self._level_ptrs(self.absSquaredGrad_all, 1)
            check(self.lib.dslam_frame_make_images(self.p, _fp(color), Bp, d, a))
            if wait:
                self.wait_host()
        else:
            check(self.lib.dslam_frame_make_images(self.p, _fp(color), Bp, None, None))

    def upload(self, color):
        color = np.ascontiguousarray(color, np.float32)
        assert color.size == self.w * self.h
        self._keep = (color, None)
        check(self.lib.dslam_frame_upload(self.p, _fp(color)))

    def build(self, B256=None):
        Bp = None
        if B256 is not None:
            B256 = np.ascontiguousarray(B256, np.float32)
            self._keep = (self._keep[0] if self._keep else None, B256)
            Bp = _fp(B256)
        check(self.lib.dslam_frame_build(self.p, Bp))

    def download(self, wait=True, levels=None, abs_grad=True):
        """D2H of the host mirrors; `levels` restricts the copy (e.g. [0] = only dI of level 0, what traceOn reads)."""
        if self.dIp_all is None:
            self.alloc_host()
        d = self._level_ptrs(self.dIp_all, 3)
        a = self._level_ptrs(self.absSquaredGrad_all, 1) if abs_grad else None
        if levels is not None:
            for l in range(self.levels):
                if l not in levels:
                    d[l] = None
                    if a is not None:
                        a[l] = None
        check(self.lib.dslam_frame_download(self.p, d, a))
        if wait:
            self.wait_host()

    def wait_host(self):
        check(self.lib.dslam_frame_wait_host(self.p))

    def dIp(self, lvl):
        wl, hl = self.sizes[lvl]
        return self.dIp_all[self.offsets[lvl]:self.offsets[lvl + 1]].reshape(hl, wl, 3)

    def absSquaredGrad(self, lvl):
        wl, hl = self.sizes[lvl]
        return self.absSquaredGrad_all[self.offsets[lvl]:self.offsets[lvl + 1]].reshape(hl, wl)


class TrackerAndScaler:
    """dso::TrackerAndScaler (src/scale_optimization/TrackerAndScaler.h:34-137) on the GPU."""

    def __init__(self, session, w, h, tfm_vec, K1, K0=None, levels=None):
        self.s = session
        self.lib = session.lib
        self.w, self.h = w, h
        self.levels = pyr_levels_used(w, h) if levels is None else levels
        K1 = np.asarray(K1, np.float32)
        if K1.shape == (3, 3):
            K1 = np.array([K1[0, 0], K1[1, 1], K1[0, 2], K1[1, 2]], np.float32)
        K0 = K1 if K0 is None else np.asarray(K0, np.float32)
        if K0.shape == (3, 3):
            K0 = np.array([K0[0, 0], K0[1, 1], K0[0, 2], K0[1, 2]], np.float32)
        K0 = np.ascontiguousarray(K0, np.float32)
        K1 = np.ascontiguousarray(K1, np.float32)
        T = np.ascontiguousarray(np.asarray(tfm_vec, np.float64).reshape(16))
        p = C.c_void_p()
        check(self.lib.dslam_ctx_create(session.p, w, h, self.levels, _fp(K0), _fp(K1), _dp(T), C.byref(p)))
        self.p = p
        # "act as pure output" members of the reference class
        self.refFrameID = -1
        self.lastRef = None
        self.lastRef_aff_g2l = (0.0, 0.0)
        self.lastFlowIndicators = np.full(3, 1000.0)
        self.firstCoarseRMSE = -1.0

    def close(self):
        if self.p:
            self.lib.dslam_ctx_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- template lifecycle ---------------------------------------------------------------------------
    def makeK(self, K0):
        K0 = np.ascontiguousarray(K0, np.float32)
        check(self.lib.dslam_ctx_make_K(self.p, _fp(K0)))

    def setAffineOptMode(self, modeA, modeB):
        check(self.lib.dslam_ctx_set_affine_mode(self.p, modeA, modeB))

    def setCoarseTrackingRef(self, ref_frame, pu, pv, pidepth, pweight, ref_aff_g2l=(0.0, 0.0), frame_id=0):
        """setCoarseTrackingRef (:317-327) with makeCoarseDepthL0 (:143-315) run on the device from the flat export of
        the active points (integer pixel, idepth, weight).  Returns pc_n per level."""
        pu = np.ascontiguousarray(pu, np.int32)
        pv = np.ascontiguousarray(pv, np.int32)
        pidepth = np.ascontiguousarray(pidepth, np.float32)
        pweight = np.ascontiguousarray(pweight, np.float32)
        pcn = np.zeros(MAX_LEVELS, np.int32)
        check(self.lib.dslam_ref_build(self.p, ref_frame.p, len(pu), _ip(pu), _ip(pv), _fp(pidepth), _fp(pweight), _ip(pcn)))
        self._set_ref_meta(ref_frame, ref_aff_g2l, frame_id)
        return pcn[:self.levels].copy()

    def setCoarseTrackingRefArrays(self, levels_uvic, ref_frame=None, ref_aff_g2l=(0.0, 0.0), frame_id=0, ref_exposure=1.0):
        """Upload host-built pc_u/pc_v/pc_idepth/pc_color per level (the first drop-in slice: makeCoarseDepthL0 stays on
        the host and only its output crosses)."""
        for lvl, (u, v, idp, col) in enumerate(levels_uvic):
            u, v, idp, col = (np.ascontiguousarray(a, np.float32) for a in (u, v, idp, col))
            check(self.lib.dslam_ref_upload(self.p, lvl, len(u), _fp(u), _fp(v), _fp(idp), _fp(col)))
        self._set_ref_meta(ref_frame, ref_aff_g2l, frame_id, ref_exposure)

    def _set_ref_meta(self, ref_frame, ref_aff_g2l, frame_id, ref_exposure=None):
        if ref_exposure is None:
            ref_exposure = ref_frame.ab_exposure if ref_frame is not None else 1.0
        check(self.lib.dslam_ref_set_affine(self.p, ref_exposure, float(ref_aff_g2l[0]), float(ref_aff_g2l[1])))
        self.lastRef = ref_frame
        self.refFrameID = frame_id
        self.lastRef_aff_g2l = (float(ref_aff_g2l[0]), float(ref_aff_g2l[1]))
        self.firstCoarseRMSE = -1.0

    def scaleCoarseDepthL0(self, scale):
        check(self.lib.dslam_ref_scale_idepth(self.p, scale))

    def ref_level(self, lvl):
        n = C.c_int(0)
        check(self.lib.dslam_ref_download(self.p, lvl, C.byref(n), None, None, None, None))
        arrs = [np.empty(n.value, np.float32) for _ in range(4)]
        check(self.lib.dslam_ref_download(self.p, lvl, C.byref(n), *[_fp(a) for a in arrs]))
        return arrs

    # ---- evaluations ------------------------------------------------------------------------------------
    def calcResAndGSPose(self, frame, lvl, pose7, aff, cutoffTH=20.0):
        """Batched calcResPose + calcGSSSEPose. pose7 [nb,7], aff [nb,2] -> dict(H [nb,8,8], b, res6, n, acc48)."""
        pose7 = np.ascontiguousarray(np.atleast_2d(pose7), np.float64)
        aff = np.ascontiguousarray(np.atleast_2d(aff), np.float64)
        nb = pose7.shape[0]
        H = np.empty((nb, 8, 8))
        b = np.empty((nb, 8))
        res = np.empty((nb, 6))
        n = np.empty(nb, np.int32)
        acc = np.empty((nb, 48))
        check(self.lib.dslam_pose_eval(self.p, frame.p, frame.ab_exposure, lvl, nb, _dp(pose7), _dp(aff), cutoffTH, _dp(H), _dp(b), _dp(res),
                                       _ip(n), _dp(acc)))
        return dict(H=H, b=b, res6=res, n=n, acc48=acc)

    def calcResAndGSScale(self, frame_right, lvl, scales, cutoffTH=20.0):
        scales = np.ascontiguousarray(np.atleast_1d(scales), np.float32)
        nb = len(scales)
        Hb = np.empty((nb, 2), np.float32)
        res = np.empty((nb, 6))
        n = np.empty(nb, np.int32)
        acc = np.empty((nb, 8))
        check(self.lib.dslam_scale_eval(self.p, frame_right.p, lvl, nb, _fp(scales), cutoffTH, _fp(Hb), _dp(res), _ip(n), _dp(acc)))
        return dict(H=Hb[:, 0].copy(), b=Hb[:, 1].copy(), res6=res, n=n, acc8=acc)

    # ---- the two reference entry points ---------------------------------------------------------------------
    def trackNewestCoarse(self, newFrameHessian, lastToNew, aff_g2l, coarsestLvl, minResForAbort=None):
        """-> (ok, lastToNew_out[7], aff_g2l_out[2], lastResiduals[5]); also sets lastFlowIndicators."""
        pose = np.array(lastToNew, np.float64).reshape(7)
        aff = np.array(aff_g2l, np.float64).reshape(2)
        mr = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
        last = np.empty(5)
        flow = np.empty(3)
        ok = C.c_int(0)
        check(self.lib.dslam_track_newest_coarse(self.p, newFrameHessian.p, newFrameHessian.ab_exposure, _dp(pose), _dp(aff), coarsestLvl,
                                                 _dp(mr), _dp(last), _dp(flow), C.byref(ok)))
        self.lastFlowIndicators = flow
        return bool(ok.value), pose, aff, last

    def trackNewestCoarseMulti(self, newFrameHessian, poses, affs, coarsestLvl, minResForAbort=None):
        poses = np.array(np.atleast_2d(poses), np.float64)
        affs = np.array(np.atleast_2d(affs), np.float64)
        nh = poses.shape[0]
        mr = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
        last = np.empty((nh, 5))
        flow = np.empty((nh, 3))
        ok = np.zeros(nh, np.int32)
        check(self.lib.dslam_track_newest_coarse_multi(self.p, newFrameHessian.p, newFrameHessian.ab_exposure, nh, _dp(poses), _dp(affs),
                                                       coarsestLvl, _dp(mr), _dp(last), _dp(flow), _ip(ok)))
        return ok.astype(bool), poses, affs, last, flow

    def trackNewCoarse(self, newFrameHessian, pose_tries, aff_init, coarsestLvl, last_coarse_rmse, reTrackThreshold=1.5):
        """The hypothesis loop of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:192-247), speculative lock-step evaluation.
        -> dict(pose, aff, achievedRes, flow, haveOneGood, tryIterations)"""
        tries = np.ascontiguousarray(np.atleast_2d(pose_tries), np.float64)
        aff0 = np.ascontiguousarray(aff_init, np.float64)
        last = np.ascontiguousarray(last_coarse_rmse, np.float64)
        pose = np.empty(7)
        aff = np.empty(2)
        ach = np.empty(5)
        flow = np.empty(3)
        good = C.c_int(0)
        ntry = C.c_int(0)
        check(self.lib.dslam_track_new_coarse(self.p, newFrameHessian.p, newFrameHessian.ab_exposure, len(tries), _dp(tries), _dp(aff0), coarsestLvl,
                                              _dp(last), reTrackThreshold, _dp(pose), _dp(aff), _dp(ach), _dp(flow), C.byref(good), C.byref(ntry)))
        return dict(pose=pose, aff=aff, achievedRes=ach, flow=flow, haveOneGood=bool(good.value), tryIterations=ntry.value)

    def optimizeScale(self, fh1, scale, coarsestLvl):
        """-> (level-0 RMSE, optimised scale)   (TrackerAndScaler.cpp:854-964)"""
        s = C.c_float(scale)
        rmse = C.c_float(0)
        check(self.lib.dslam_optimize_scale(self.p, fh1.p, C.byref(s), coarsestLvl, C.byref(rmse)))
        return rmse.value, s.value

    def optimizeScaleMulti(self, fh1, scales, coarsestLvl):
        """The seed loop of FrontEnd::optimizeScale (src/FrontEnd.cpp:995-1003) in lock step."""
        scales = np.array(np.atleast_1d(scales), np.float32)
        rmse = np.empty(len(scales), np.float32)
        check(self.lib.dslam_optimize_scale_multi(self.p, fh1.p, len(scales), _fp(scales), coarsestLvl, _fp(rmse)))
        return rmse, scales

    def trace(self):
        n = C.c_int(0)
        check(self.lib.dslam_get_trace(self.p, None, 0, C.byref(n)))
        out = np.zeros((n.value, 15))
        if n.value:
            check(self.lib.dslam_get_trace(self.p, _dp(out), n.value, C.byref(n)))
        return out

    def counters(self):
        out = (C.c_longlong * 3)()
        check(self.lib.dslam_ctx_counters(self.p, out))
        return dict(evals=out[0], launches=out[1], iterations=out[2])


class PoseEstimator:
    """dso::PoseEstimator (src/loop_closure/pose_estimation/PoseEstimator.h:34-49) on the GPU: direct alignment of a
    loop-closure candidate (LoopHandler.cpp:274-277)."""

    def __init__(self, session, w, h, levels=None):
        self.s = session
        self.lib = session.lib
        self.w, self.h = w, h
        self.levels = pyr_levels_used(w, h) if levels is None else levels
        p = C.c_void_p()
        check(self.lib.dslam_pe_create(session.p, w, h, self.levels, C.byref(p)))
        self.p = p
        self.inlier_percent = 0

    def close(self):
        if self.p:
            self.lib.dslam_pe_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setAffineOptMode(self, modeA, modeB):
        check(self.lib.dslam_pe_set_affine_mode(self.p, modeA, modeB))

    def setPoints(self, pts, colors, ref_ab_exposure):
        """pts [n,3] float64 (pair.first), colors [n, levels] float32 (pair.second[lvl])."""
        pts = np.ascontiguousarray(pts, np.float64)
        colors = np.ascontiguousarray(colors, np.float32)
        if pts.ndim != 2 or pts.shape[1] != 3 or colors.shape != (len(pts), self.levels):
            raise ValueError("pts must be [n,3] and colors [n,levels]")
        check(self.lib.dslam_pe_set_points(self.p, len(pts), _dp(pts), _fp(colors), ref_ab_exposure))

    def calcResAndGS(self, new_fh, new_cam, lvl, ref_to_new, aff=(0.0, 0.0), cutoffTH=20.0):
        cam = np.ascontiguousarray(new_cam, np.float32)
        T = np.ascontiguousarray(np.asarray(ref_to_new, np.float64).reshape(16))
        aff = np.ascontiguousarray(aff, np.float64)
        H, b, res, acc = np.empty((8, 8)), np.empty(8), np.empty(6), np.empty(48)
        n = np.empty(1, np.int32)
        check(self.lib.dslam_pe_eval(self.p, new_fh.p, new_fh.ab_exposure, _fp(cam), lvl, _dp(T), _dp(aff), cutoffTH, _dp(H), _dp(b), _dp(res),
                                     _ip(n), _dp(acc)))
        return dict(H=H, b=b, res6=res, n=int(n[0]), acc48=acc)

    def estimate(self, pts, ref_ab_exposure, new_fh, new_cam, coarsest_lvl, ref_to_new):
        """bool estimate(pts, ref_ab_exposure, new_fh, new_cam, coarsest_lvl, ref_to_new&, pose_error&) (:298-304).
        pts = (xyz [n,3], colors [n,levels]) or None to reuse the last setPoints.  Returns (ok, ref_to_new 4x4, pose_error)."""
        if pts is not None:
            self.setPoints(pts[0], pts[1], ref_ab_exposure)
        cam = np.ascontiguousarray(new_cam, np.float32)
        T = np.array(ref_to_new, np.float64).reshape(16)
        err = C.c_float(0)
        inl = C.c_int(0)
        ok = C.c_int(0)
        check(self.lib.dslam_pe_estimate(self.p, new_fh.p, new_fh.ab_exposure, _fp(cam), coarsest_lvl, _dp(T), C.byref(err), C.byref(inl),
                                         C.byref(ok)))
        self.inlier_percent = inl.value
        return bool(ok.value), T.reshape(4, 4), err.value

    def trace(self):
        n = C.c_int(0)
        check(self.lib.dslam_pe_get_trace(self.p, None, 0, C.byref(n)))
        out = np.zeros((n.value, 15))
        if n.value:
            check(self.lib.dslam_pe_get_trace(self.p, _dp(out), n.value, C.byref(n)))
        return out


class ScanContextDB:
    """Scan-Context descriptor database; replaces search_ringkey / search_sc (search_place.h:25-85)."""

    KEY_EMPTY = 0xFFFFFFFFFFFFFFFF

    def __init__(self, session, capacity, n_sectors=60, n_rings=20, fp64=False):
        """capacity = initial allocation (the tables grow); fp64 keeps the reference's double signature values next to the
        fp32 scan table so that the exact re-score multiplies doubles like search_place.h:71-77."""
        self.s = session
        self.lib = session.lib
        self.n_sectors, self.n_rings = n_sectors, n_rings
        self.n_cells = n_sectors * n_rings
        self.fp64 = bool(fp64)
        p = C.c_void_p()
        check(self.lib.dslam_sc_create_ex(session.p, n_sectors, n_rings, capacity, 1 if fp64 else 0, C.byref(p)))
        self.p = p
        self.world, self.rank = 1, 0

    def close(self):
        if self.p:
            self.lib.dslam_sc_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        n = C.c_int(0)
        check(self.lib.dslam_sc_size(self.p, C.byref(n)))
        return n.value

    def add(self, ringkeys, sigs, global_ids=None):
        ringkeys = np.ascontiguousarray(np.atleast_2d(ringkeys), np.float32)
        sigs = np.ascontiguousarray(np.atleast_2d(sigs), np.float32)
        assert ringkeys.shape[1] == self.n_rings and sigs.shape[1] == self.n_cells and len(sigs) == len(ringkeys)
        ids = None if global_ids is None else np.ascontiguousarray(global_ids, np.int32)
        check(self.lib.dslam_sc_add(self.p, len(sigs), _fp(ringkeys), _fp(sigs), _ip(ids)))

    def add64(self, ringkeys, sigs64, global_ids=None):
        ringkeys = np.ascontiguousarray(np.atleast_2d(ringkeys), np.float32)
        sigs64 = np.ascontiguousarray(np.atleast_2d(sigs64), np.float64)
        assert ringkeys.shape[1] == self.n_rings and sigs64.shape[1] == self.n_cells and len(sigs64) == len(ringkeys)
        ids = None if global_ids is None else np.ascontiguousarray(global_ids, np.int32)
        check(self.lib.dslam_sc_add64(self.p, len(sigs64), _fp(ringkeys), _dp(sigs64), _ip(ids)))

    def save(self, path):
        """Dump the shard as a .scdb file (format: include/dslam_b200.h)."""
        check(self.lib.dslam_sc_save(self.p, str(path).encode()))

    def load(self, path):
        """Append the rows of a .scdb file; returns the number of rows read."""
        n = C.c_int(0)
        check(self.lib.dslam_sc_load(self.p, str(path).encode(), C.byref(n)))
        return n.value

    def add_sparse(self, ringkey, idx, val, global_id=-1):
        ringkey = np.ascontiguousarray(ringkey, np.float32)
        idx = np.ascontiguousarray(idx, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        check(self.lib.dslam_sc_add_sparse(self.p, _fp(ringkey), _ip(idx), _dp(val), len(idx), global_id))

    def generate(self, pts_spherical, lidar_range=40.0, append=False, global_id=-1):
        """ScanContext::generate (ScanContext.cpp:78-142) on the device -> (ringkey [r], dense signature fp32 [s*r], fp64 values, tfm_pca_rig 4x4)."""
        pts = np.ascontiguousarray(pts_spherical, np.float64).reshape(-1, 3)
        rk = np.empty(self.n_rings, np.float32)
        sig = np.empty(self.n_cells, np.float32)
        sig64 = np.empty(self.n_cells, np.float64)
        tfm = np.empty(16, np.float64)
        check(self.lib.dslam_sc_generate(self.p, _dp(pts), len(pts), lidar_range, _fp(rk), _fp(sig), _dp(sig64), _dp(tfm), 1 if append else 0, global_id))
        return rk, sig, sig64, tfm.reshape(4, 4)

    def search_ringkey(self, ringkeys, k=3, thres=0.1, max_id=2**31 - 1):
        ringkeys = np.ascontiguousarray(np.atleast_2d(ringkeys), np.float32)
        nq = len(ringkeys)
        cand = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.float32)
        check(self.lib.dslam_sc_search_ringkey(self.p, nq, _fp(ringkeys), k, thres, max_id, _ip(cand), _fp(dist)))
        return cand, dist

    def search_sc(self, sigs, candidates):
        sigs = np.ascontiguousarray(np.atleast_2d(sigs), np.float32)
        candidates = np.ascontiguousarray(np.atleast_2d(candidates), np.int32)
        nq = len(sigs)
        idx = np.empty(nq, np.int32)
        diff = np.empty(nq, np.float32)
        check(self.lib.dslam_sc_search_sc(self.p, nq, _fp(sigs), _ip(candidates), candidates.shape[1], _ip(idx), _fp(diff)))
        return idx, diff

    def search_sc64(self, sigs64, candidates):
        sigs64 = np.ascontiguousarray(np.atleast_2d(sigs64), np.float64)
        candidates = np.ascontiguousarray(np.atleast_2d(candidates), np.int32)
        nq = len(sigs64)
        idx = np.empty(nq, np.int32)
        diff = np.empty(nq, np.float32)
        check(self.lib.dslam_sc_search_sc64(self.p, nq, _dp(sigs64), _ip(candidates), candidates.shape[1], _ip(idx), _fp(diff)))
        return idx, diff

    def query(self, sigs, ringkeys=None, ringkey_thres=-1.0, max_id=2**31 - 1):
        sigs = np.ascontiguousarray(np.atleast_2d(sigs), np.float32)
        nq = len(sigs)
        rk = None if ringkeys is None else np.ascontiguousarray(np.atleast_2d(ringkeys), np.float32)
        idx = np.empty(nq, np.int32)
        diff = np.empty(nq, np.float32)
        check(self.lib.dslam_sc_query(self.p, nq, _fp(rk), _fp(sigs), ringkey_thres, max_id, _ip(idx), _fp(diff)))
        return idx, diff

    def query64(self, sigs64, ringkeys=None, ringkey_thres=-1.0, max_id=2**31 - 1):
        sigs64 = np.ascontiguousarray(np.atleast_2d(sigs64), np.float64)
        nq = len(sigs64)
        rk = None if ringkeys is None else np.ascontiguousarray(np.atleast_2d(ringkeys), np.float32)
        idx = np.empty(nq, np.int32)
        diff = np.empty(nq, np.float32)
        check(self.lib.dslam_sc_query64(self.p, nq, _fp(rk), _dp(sigs64), ringkey_thres, max_id, _ip(idx), _fp(diff)))
        return idx, diff

    def exchange_mode(self):
        """'nvlink-mailbox' when the sharded query combines its keys through peer memory, 'nccl' for the all-reduce fallback."""
        m = C.c_int(0)
        check(self.lib.dslam_sc_exchange_mode(self.p, C.byref(m)))
        return "nvlink-mailbox" if m.value else ("nccl" if self.world > 1 else "none")

    def query_keys(self, sigs, ringkeys=None, ringkey_thres=-1.0, max_id=2**31 - 1):
        sigs = np.ascontiguousarray(np.atleast_2d(sigs), np.float32)
        nq = len(sigs)
        rk = None if ringkeys is None else np.ascontiguousarray(np.atleast_2d(ringkeys), np.float32)
        keys = np.empty(nq, np.uint64)
        check(self.lib.dslam_sc_query_keys(self.p, nq, _fp(rk), _fp(sigs), ringkey_thres, max_id, keys.ctypes.data_as(c_u64)))
        return keys

    def last_scan_ms(self):
        ms = C.c_float(0)
        check(self.lib.dslam_sc_last_scan_ms(self.p, C.byref(ms)))
        return ms.value

    def set_scan_kernel(self, flavour):
        """auto / stream / tile / umma (tcgen05) / umma_masked (process-wide; dslam_sc_set_scan_kernel)."""
        check(self.lib.dslam_sc_set_scan_kernel({"auto": 0, "stream": 1, "tile": 2, "umma": 3, "umma_masked": 4}.get(flavour, flavour)))

    def attach_comm(self, id128, world_size, rank):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(id128))
        check(self.lib.dslam_sc_comm_init(self.p, buf, world_size, rank))
        self.world, self.rank = world_size, rank

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        check(_lib.load().dslam_sc_unique_id(buf))
        return bytes(buf)


# ---- packed (distance, id) keys: the value a min-reduction over shards combines ------------------------------------
def pack_key(dist, idx):
    """Order-preserving packing used by the device kernels: total-order bits of the fp32 distance << 32 | id."""
    b = np.asarray(dist, np.float32).view(np.uint32).astype(np.uint64)
    neg = (b >> np.uint64(31)) != 0
    b = np.where(neg, b ^ np.uint64(0xFFFFFFFF), b ^ np.uint64(0x80000000))
    return (b << np.uint64(32)) | np.asarray(idx, np.int64).astype(np.uint64)


def unpack_key(key):
    key = np.asarray(key, np.uint64)
    empty = key == np.uint64(ScanContextDB.KEY_EMPTY)
    b = (key >> np.uint64(32)).astype(np.uint32)
    neg = (b >> np.uint32(31)) == 0
    b = np.where(neg, b ^ np.uint32(0xFFFFFFFF), b ^ np.uint32(0x80000000)).astype(np.uint32)
    dist = b.view(np.float32)
    idx = (key & np.uint64(0xFFFFFFFF)).astype(np.int64).astype(np.int32)
    return np.where(empty, np.float32(1.1), dist), np.where(empty, -1, idx)


def shard_rows(n_rows, world_size, rank):
    """Row i of the global database lives on rank i % world_size (ids stay ascending inside every shard)."""
    return np.arange(rank, n_rows, world_size, dtype=np.int64)


def combine_keys_torch(keys, group=None):
    """min+argmin across ranks with torch.distributed on packed keys (works on any backend; gloo in the CPU tests).
    uint64 order is mapped onto int64 by flipping the top bit."""
    import torch
    import torch.distributed as dist

    k = torch.from_numpy((np.asarray(keys, np.uint64) ^ np.uint64(1 << 63)).view(np.int64).copy())
    dist.all_reduce(k, op=dist.ReduceOp.MIN, group=group)
    return k.numpy().view(np.uint64) ^ np.uint64(1 << 63)


# ---- independent stereo streams in lock step (one kernel launch per LM round for all of them) ---------------------
def track_newest_coarse_batch(trackers, frames, poses, affs, coarsestLvl, minResForAbort=None):
    n = len(trackers)
    lib = trackers[0].lib
    ctxs = (C.c_void_p * n)(*[t.p for t in trackers])
    frs = (C.c_void_p * n)(*[f.p for f in frames])
    expo = np.ascontiguousarray([f.ab_exposure for f in frames], np.float32)
    poses = np.array(poses, np.float64).reshape(n, 7)
    affs = np.array(affs, np.float64).reshape(n, 2)
    mr = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
    last = np.empty((n, 5))
    flow = np.empty((n, 3))
    ok = np.zeros(n, np.int32)
    check(lib.dslam_track_newest_coarse_batch(n, ctxs, frs, _fp(expo), _dp(poses), _dp(affs), coarsestLvl, _dp(mr), _dp(last), _dp(flow), _ip(ok)))
    for i, t in enumerate(trackers):
        t.lastFlowIndicators = flow[i]
    return ok.astype(bool), poses, affs, last


def optimize_scale_batch(trackers, frames_right, scales, coarsestLvl):
    n = len(trackers)
    lib = trackers[0].lib
    ctxs = (C.c_void_p * n)(*[t.p for t in trackers])
    frs = (C.c_void_p * n)(*[f.p for f in frames_right])
    scales = np.array(scales, np.float32).reshape(n)
    rmse = np.empty(n, np.float32)
    check(lib.dslam_optimize_scale_batch(n, ctxs, frs, _fp(scales), coarsestLvl, _fp(rmse)))
    return rmse, scales


def lm_batch(pose_trackers, pose_frames, poses, affs, coarsestLvl, scale_trackers, scale_frames, scales, scale_coarsestLvl=None,
             minResForAbort=None):
    """n tracking jobs and m scale-optimisation jobs in ONE lock step (dslam_lm_batch): the pose and the scale LM loops are
    independent, so every round of all of them is one launch.  Returns (ok, poses, affs, lastResiduals, rmse, scales)."""
    n, m = len(pose_trackers), len(scale_trackers)
    lib = (pose_trackers or scale_trackers)[0].lib
    pc = (C.c_void_p * max(n, 1))(*[t.p for t in pose_trackers])
    pf = (C.c_void_p * max(n, 1))(*[f.p for f in pose_frames])
    sc = (C.c_void_p * max(m, 1))(*[t.p for t in scale_trackers])
    sf = (C.c_void_p * max(m, 1))(*[f.p for f in scale_frames])
    expo = np.ascontiguousarray([f.ab_exposure for f in pose_frames] or [1.0], np.float32)
    poses = np.array(poses, np.float64).reshape(max(n, 0), 7) if n else np.zeros((0, 7))
    affs = np.array(affs, np.float64).reshape(n, 2) if n else np.zeros((0, 2))
    mr = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
    last = np.empty((n, 5))
    flow = np.empty((n, 3))
    ok = np.zeros(max(n, 1), np.int32)
    scales = np.array(scales, np.float32).reshape(m) if m else np.zeros(0, np.float32)
    rmse = np.empty(max(m, 1), np.float32)
    check(lib.dslam_lm_batch(n, pc, pf, _fp(expo), _dp(poses) if n else None, _dp(affs) if n else None, coarsestLvl, _dp(mr), _dp(last) if n else None,
                             _dp(flow) if n else None, _ip(ok), m, sc, sf, _fp(scales) if m else None,
                             coarsestLvl if scale_coarsestLvl is None else scale_coarsestLvl, _fp(rmse)))
    for i, t in enumerate(pose_trackers):
        t.lastFlowIndicators = flow[i]
    return ok[:n].astype(bool), poses, affs, last, rmse[:m], scales


class LmBatchPlan:
    """A fixed set of tracking + scale jobs that is run again and again (the streams of a capture rig): the pointer arrays and
    result buffers of lm_batch are built once.  run() = dslam_lm_batch; same results as lm_batch()."""

    def __init__(self, pose_trackers, pose_frames, scale_trackers, scale_frames, coarsestLvl, minResForAbort=None):
        self.n, self.m = len(pose_trackers), len(scale_trackers)
        n, m = self.n, self.m
        self.lib = (pose_trackers or scale_trackers)[0].lib
        self.pose_trackers = list(pose_trackers)
        self._keep = (list(pose_frames), list(scale_trackers), list(scale_frames))
        self.pc = (C.c_void_p * max(n, 1))(*[t.p for t in pose_trackers])
        self.pf = (C.c_void_p * max(n, 1))(*[f.p for f in pose_frames])
        self.sc = (C.c_void_p * max(m, 1))(*[t.p for t in scale_trackers])
        self.sf = (C.c_void_p * max(m, 1))(*[f.p for f in scale_frames])
        self.expo = np.ascontiguousarray([f.ab_exposure for f in pose_frames] or [1.0], np.float32)
        self.mr = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
        self.lvl = coarsestLvl
        self.poses = np.zeros((n, 7))
        self.affs = np.zeros((n, 2))
        self.last = np.empty((n, 5))
        self.flow = np.empty((n, 3))
        self.ok = np.zeros(max(n, 1), np.int32)
        self.scales = np.zeros(m, np.float32)
        self.rmse = np.empty(max(m, 1), np.float32)

    def run(self, poses, affs, scales):
        n, m = self.n, self.m
        if n:
            self.poses[...] = poses
            self.affs[...] = affs
        if m:
            self.scales[...] = scales
        check(self.lib.dslam_lm_batch(n, self.pc, self.pf, _fp(self.expo), _dp(self.poses) if n else None, _dp(self.affs) if n else None, self.lvl,
                                      _dp(self.mr), _dp(self.last) if n else None, _dp(self.flow) if n else None, _ip(self.ok), m, self.sc, self.sf,
                                      _fp(self.scales) if m else None, self.lvl, _fp(self.rmse)))
        for i, t in enumerate(self.pose_trackers):
            t.lastFlowIndicators = self.flow[i]
        return self.ok[:n].astype(bool), self.poses, self.affs, self.last, self.rmse[:m], self.scales


class FrameBatchPlan:
    """A fixed set of frames whose images arrive from fixed host buffers: pointer arrays built once.
    upload() = dslam_frame_upload_batch, build() = dslam_frame_build_batch."""

    def __init__(self, frames, images=None):
        self.frames = list(frames)
        self.lib = frames[0].lib
        self.n = len(frames)
        self.fa = (C.c_void_p * self.n)(*[f.p for f in frames])
        self.images = None
        if images is not None:
            self.images = [np.ascontiguousarray(im, np.float32) for im in images]
            for f, im in zip(frames, self.images):
                assert im.size == f.w * f.h
            self.ia = (C.c_void_p * self.n)(*[im.ctypes.data for im in self.images])

    def upload(self):
        check(self.lib.dslam_frame_upload_batch(self.n, self.fa, self.ia))

    def build(self, stage_host=0, overlap=False):
        check(self.lib.dslam_frame_build_batch(self.n, self.fa, None, stage_host | (4 if overlap else 0)))


def upload_frames(frames, images):
    """Raw images of n frames (dslam_frame_upload_batch): images that lie back to back in host memory go up as one transfer."""
    n = len(frames)
    images = [np.ascontiguousarray(im, np.float32) for im in images]
    for f, im in zip(frames, images):
        assert im.size == f.w * f.h
        f._keep = (im, None)
    fa = (C.c_void_p * n)(*[f.p for f in frames])
    ia = (C.c_void_p * n)(*[im.ctypes.data for im in images])
    check(frames[0].lib.dslam_frame_upload_batch(n, fa, ia))


def build_frames(frames, B256=None, stage_host=0, overlap=False):
    """Pyramids of all (uploaded) frames in two kernel launches (dslam_frame_build_batch).  overlap=True builds them on the
    session's pyramid stream, concurrently with tracking calls on other frames queued afterwards (stage_host bit 2)."""
    if overlap:
        stage_host |= 4
    n = len(frames)
    arr = (C.c_void_p * n)(*[f.p for f in frames])
    Bp = None
    if B256 is not None:
        B256 = np.ascontiguousarray(B256, np.float32)
        Bp = _fp(B256)
    check(frames[0].lib.dslam_frame_build_batch(n, arr, Bp, stage_host))
