"""The hand-written oracle (oracle/dslam_oracle.cpp) against the REFERENCE'S OWN TrackerAndScaler.cpp hot path compiled in
place (oracle/ref_build.py -> oracle/_ref/libdslam_ref_tracker.so: lines 1-336 and 451-1172 of the reference file + DSO's
real MatrixAccumulators.h / globalFuncs.h + the reference's ScaleAccumulator.h, against the Eigen / Sophus / DSO-struct
stand-ins of oracle/shim).  Everything is compared BIT FOR BIT in the oracle's reference-faithful mode (0: fp32, SSE order).

This pins every formula and the whole control flow of the path to the reference's source text.  What remains unpinned is
only what the shim itself restates: the evaluation order real Eigen gives the small fixed-size expressions and Sophus'
SE3 exp / product (documented in oracle/shim/Eigen/Core and DESIGN.md)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as orc
from helpers import IDENT7, OracleCase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def have_ref():
    if not orc.ReferenceTracker.available():
        if os.path.isdir("/root/reference"):
            subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_build.py")], check=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference not present")
    return True


def make_pair(oracle, cfg, seed, hdif_spread=True, **kw):
    oc = OracleCase(oracle, cfg, seed, **kw)
    c = oc.case
    rng = np.random.default_rng(seed)
    hdif = (rng.uniform(2e-4, 5e-3, len(c["pu"])) if hdif_spread else np.full(len(c["pu"]), 1e-3)).astype(np.float32)
    w = orc.ReferenceTracker.weight_from_hdif(hdif)
    oc.trk.make_coarse_depth(c["pu"], c["pv"], c["pid"], w, oc.dIp_ref)
    rt = orc.ReferenceTracker(oc.w, oc.h, oc.levels, oc.K, oc.K, oc.T_stereo)
    rt.set_ref(oc.dIp_ref, c["pu"], c["pv"], c["pid"], hdif)
    rt.set_new_frame(oc.dIp_new)
    rt.set_right_frame(oc.dIp_right)
    return oc, rt


@pytest.mark.parametrize("seed", [3, 5, 12])
def test_template_and_intrinsics(oracle, have_ref, seed):
    """ctor / makeK / makeCoarseDepthL0 (:47-141, 143-315): pc_* arrays, Ki, level intrinsics."""
    oc, rt = make_pair(oracle, "tiny", seed)
    for l in range(oc.levels):
        for a, b in zip(rt.get_ref_level(l), oc.trk.get_ref_level(l)):
            assert np.array_equal(bits(a), bits(b))
        k, ko = rt.get_K(l), oc.trk.get_K(l)
        assert np.array_equal(bits(k[:4]), bits([ko["fx"], ko["fy"], ko["cx"], ko["cy"]]))
        assert np.array_equal(bits(k[4:13]), bits(ko["Ki"].reshape(-1)))
        assert np.array_equal(bits(k[13:]), bits([ko["fx1"], ko["fy1"], ko["cx1"], ko["cy1"]]))
    rt.scale_idepth(1.37)
    oc.trk.scale_idepth(1.37)
    for l in range(oc.levels):
        assert np.array_equal(bits(rt.get_ref_level(l)[2]), bits(oc.trk.get_ref_level(l)[2]))


def test_calc_res_and_gs_pose(oracle, have_ref):
    """calcResPose (:699-852) incl. the warped buffers and flow indicators, calcGSSSEPose (:640-697)."""
    oc, rt = make_pair(oracle, "tiny", 3)
    rng = np.random.default_rng(0)
    poses = [IDENT7, oc.case["pose7_true"]] + [oracle.se3_mul(oracle.se3_exp(np.concatenate([rng.normal(0, 0.02, 3), rng.normal(0, 0.004, 3)])),
                                                              oc.case["pose7_true"]) for _ in range(4)]
    affs = [(0.0, 0.0), (0.03, 4.0)] + [tuple(rng.normal(0, [0.05, 5.0])) for _ in range(4)]
    for lvl in range(oc.levels):
        for pose, aff, cutoff in zip(poses, affs, (20.0, 20.0, 40.0, 5.0, 160.0, 20.0)):
            r1, n1 = rt.calc_res_pose(lvl, pose, aff, cutoff)
            r2, n2 = oc.trk.calc_res_pose(lvl, pose, aff, cutoff)
            assert n1 == n2 and np.array_equal(r1, r2, equal_nan=True)
            assert np.array_equal(rt.get_warped(0).view(np.uint32), oc.trk.get_warped(0).view(np.uint32))
            H1, b1 = rt.calc_gs_pose(lvl, pose, aff)
            H2, b2, _ = oc.trk.calc_gs_pose(lvl, 0, aff)
            assert np.array_equal(H1, H2) and np.array_equal(b1, b2)


def test_calc_res_and_gs_scale(oracle, have_ref):
    """calcResScale (:1007-1172), calcGSSSEScale (:966-1005)."""
    oc, rt = make_pair(oracle, "tiny", 4, scale_error=2.5)
    for lvl in range(oc.levels):
        for s, cutoff in ((1.0, 20.0), (0.3, 20.0), (2.5, 40.0), (12.0, 160.0)):
            r1, n1 = rt.calc_res_scale(lvl, s, cutoff)
            r2, n2 = oc.trk.calc_res_scale(lvl, s, cutoff)
            assert n1 == n2 and np.array_equal(r1, r2, equal_nan=True)
            assert np.array_equal(rt.get_warped(1).view(np.uint32), oc.trk.get_warped(1).view(np.uint32))
            H1, b1 = rt.calc_gs_scale(lvl, s)
            H2, b2, _ = oc.trk.calc_gs_scale(lvl, 0, s)
            assert np.float32(H1) == np.float32(H2) and np.float32(b1) == np.float32(b2)


@pytest.mark.parametrize("seed,motion", [(3, 1.0), (5, 0.5), (8, 1.5), (12, 2.5)])
def test_track_newest_coarse(oracle, have_ref, seed, motion):
    """trackNewestCoarse (:451-638): pose, affine, per-level residuals and flow indicators, bit for bit."""
    oc, rt = make_pair(oracle, "tiny", seed, motion_scale=motion)
    for start in (IDENT7, oc.case["pose7_true"]):
        ok1, p1, a1, l1, f1 = rt.track_newest_coarse(start, (0, 0), oc.levels - 1)
        ok2, p2, a2, l2, f2 = oc.trk.track_newest_coarse(0, start, (0, 0), oc.levels - 1)
        assert ok1 == ok2
        assert np.array_equal(p1, p2) and np.array_equal(a1, a2)
        assert np.array_equal(l1, l2, equal_nan=True) and np.array_equal(f1, f2)


def test_track_abort_and_partial_pyramid(oracle, have_ref):
    oc, rt = make_pair(oracle, "tiny", 3)
    r1 = rt.track_newest_coarse(IDENT7, (0, 0), oc.levels - 1, min_res=np.full(5, 0.1))
    r2 = oc.trk.track_newest_coarse(0, IDENT7, (0, 0), oc.levels - 1, min_res=np.full(5, 0.1))
    assert r1[0] == r2[0] == False  # noqa: E712  aborted on the coarsest level
    assert np.array_equal(r1[1], IDENT7) and np.array_equal(r2[1], IDENT7)  # outputs untouched (:612-613 not reached)
    assert np.array_equal(r1[3], r2[3], equal_nan=True)
    r1 = rt.track_newest_coarse(IDENT7, (0, 0), 1)  # coarsestLvl below the top of the pyramid
    r2 = oc.trk.track_newest_coarse(0, IDENT7, (0, 0), 1)
    assert r1[0] == r2[0] and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[3], r2[3], equal_nan=True)


@pytest.mark.parametrize("modeA,modeB", [(-1, -1), (0, -1), (-1, 0), (1, 1)])
def test_track_affine_modes(oracle, have_ref, modeA, modeB):
    """setting_affineOptModeA/B < 0 fix a and / or b (:511-534), != 0 enable the plausibility limits (:615-626)."""
    oc, rt = make_pair(oracle, "tiny", 6)
    try:
        rt.set_aff_mode(modeA, modeB)
        oc.trk.set_aff_mode(modeA, modeB)
        ok1, p1, a1, l1, f1 = rt.track_newest_coarse(IDENT7, (0, 0), oc.levels - 1)
        ok2, p2, a2, l2, f2 = oc.trk.track_newest_coarse(0, IDENT7, (0, 0), oc.levels - 1)
        assert ok1 == ok2 and np.array_equal(p1, p2) and np.array_equal(a1, a2) and np.array_equal(l1, l2, equal_nan=True)
    finally:
        rt.set_aff_mode(0, 0)  # the reference keeps these in process-wide globals


@pytest.mark.parametrize("seed,scale_error,seeds", [(4, 2.5, (0.1, 1, 5, 10, 15, 25, 30, 50)), (7, 0.4, (1.0, 0.2)), (9, 1.0, (1.0, 3.0))])
def test_optimize_scale(oracle, have_ref, seed, scale_error, seeds):
    """optimizeScale (:854-964) from the seeds of FrontEnd::optimizeScale (src/FrontEnd.cpp:995-1003)."""
    oc, rt = make_pair(oracle, "tiny", seed, scale_error=scale_error)
    for s0 in seeds:
        rm1, s1 = rt.optimize_scale(s0, oc.levels - 1)
        rm2, s2 = oc.trk.optimize_scale(0, s0, oc.levels - 1)
        assert np.float32(s1) == np.float32(s2)
        assert np.float32(rm1) == np.float32(rm2) or (np.isnan(rm1) and np.isnan(rm2))


@pytest.mark.parametrize("w,h,levels,gamma", [(320, 192, 3, False), (154, 46, 2, True), (1232, 368, 5, True), (150, 94, 2, False)])
def test_make_images(oracle, have_ref, w, h, levels, gamma):
    """FrameHessian::makeImages (deps:dso/src/FullSystem/HessianBlocks.cpp:128-191) compiled from the reference's text."""
    rng = np.random.default_rng(w + h)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.clip(128 + 60 * np.sin(xx * 0.11) * np.cos(yy * 0.07) + rng.normal(0, 8, (h, w)), 0, 255).astype(np.float32)
    B = (255.0 * (np.arange(256) / 255.0) ** 0.8).astype(np.float32) if gamma else None
    d_ref, a_ref = orc.reference_make_images(img, levels, B)
    d_orc, a_orc = oracle.make_images(img, levels, B)
    assert np.array_equal(d_ref.view(np.uint32), d_orc.view(np.uint32))
    assert np.array_equal(a_ref.view(np.uint32), a_orc.view(np.uint32))
