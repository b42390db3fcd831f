"""The oracle's PoseEstimator restatement against the REFERENCE'S OWN src/loop_closure/pose_estimation/PoseEstimator.cpp
compiled in place (oracle/ref_build.py -> oracle/_ref/libdslam_ref_pe.so): bit for bit in the reference-faithful mode."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as orc
from helpers import OracleCase, loop_closure_points, mat4_from_pose7

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def have_ref():
    if not orc.ReferencePoseEstimator.available():
        if os.path.isdir("/root/reference"):
            subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_build.py")], check=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference not present")
    return True


def make(oracle, seed, motion=1.0):
    oc = OracleCase(oracle, "tiny", seed, motion_scale=motion)
    pts, colors = loop_closure_points(oc, 1500, seed)
    pe = oracle.pose_estimator(oc.w, oc.h, oc.levels, oc.K)
    rp = orc.ReferencePoseEstimator(oc.w, oc.h, oc.levels, oc.K)
    for x in (pe, rp):
        x.set_points(pts, colors, 1.0)
        x.set_new_frame(oc.dIp_new, 1.0)
    return oc, pe, rp


def test_calc_res_and_gs(oracle, have_ref):
    oc, pe, rp = make(oracle, 3)
    rng = np.random.default_rng(1)
    T_true = mat4_from_pose7(oracle, oc.case["pose7_true"])
    for lvl in range(oc.levels):
        for T, aff, cutoff in ((np.eye(4), (0, 0), 20.0), (T_true, (0.03, 4.0), 20.0), (T_true, (0.0, -3.0), 40.0)):
            r1, n1, H1, b1 = rp.calc_res(lvl, T, aff, cutoff)
            r2, n2, H2, b2, _ = pe.calc_res(lvl, 0, T, aff, cutoff)
            assert n1 == n2 and np.array_equal(r1, r2, equal_nan=True)
            assert np.array_equal(H1, H2) and np.array_equal(b1, b2)


@pytest.mark.parametrize("seed,motion", [(3, 1.0), (5, 0.5), (8, 1.5)])
def test_estimate(oracle, have_ref, seed, motion):
    oc, pe, rp = make(oracle, seed, motion)
    for T0 in (np.eye(4), mat4_from_pose7(oracle, oc.case["pose7_true"])):
        ok1, T1, e1 = rp.estimate(T0, oc.levels - 1)
        ok2, T2, e2, inl = pe.estimate(0, T0, oc.levels - 1)
        assert ok1 == ok2 and np.array_equal(T1, T2) and np.float32(e1) == np.float32(e2)
    # the alignment recovers the true motion
    assert np.linalg.norm(T2[:3, 3] - oc.case["t_true"]) < 0.03 * np.linalg.norm(oc.case["t_true"]) + 2e-3


@pytest.mark.parametrize("modeA,modeB", [(-1, -1), (0, -1), (-1, 0), (1, 1)])
def test_estimate_affine_modes(oracle, have_ref, modeA, modeB):
    oc, pe, rp = make(oracle, 6)
    try:
        rp.set_aff_mode(modeA, modeB)
        pe.set_aff_mode(modeA, modeB)
        ok1, T1, e1 = rp.estimate(np.eye(4), oc.levels - 1)
        ok2, T2, e2, _ = pe.estimate(0, np.eye(4), oc.levels - 1)
        assert ok1 == ok2 and np.array_equal(T1, T2) and np.float32(e1) == np.float32(e2)
    finally:
        rp.set_aff_mode(0, 0)
