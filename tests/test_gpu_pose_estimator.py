"""PoseEstimator (loop-closure direct alignment) on the GPU vs the CPU oracle, which is itself
bit-identical to the reference's PoseEstimator.cpp compiled in place (tests/test_oracle_ref_pe.py).

Tolerances as in test_gpu_tracker.py: counts exact; raw sums 1e-11 relative against the oracle's fp64-accumulating mode;
per-iteration LM increments 1e-5 relative with identical accept / reject sequences; the boolean result, inlier percentage
and number of iterations exact.
"""
import numpy as np
import pytest

from helpers import OracleCase, loop_closure_points, mat4_from_pose7, rel_err

pytestmark = pytest.mark.gpu

TOL_SUM = 1e-11
TOL_INC = 1e-5


class PeCase:
    def __init__(self, session, oracle, seed, motion=1.0, n=1500, cfg="tiny"):
        from direct_stereo_slam_b200 import api

        self.oc = oc = OracleCase(oracle, cfg, seed, motion_scale=motion)
        self.pts, self.colors = loop_closure_points(oc, n, seed)  # colors level-major
        self.ope = oracle.pose_estimator(oc.w, oc.h, oc.levels, oc.K)
        self.ope.set_points(self.pts, self.colors, 1.0)
        self.ope.set_new_frame(oc.dIp_new, 1.0)
        self.f_new = api.FrameHessian(session, oc.w, oc.h, oc.levels)
        self.f_new.makeImages(oc.case["img_new"], host=False)
        self.gpe = api.PoseEstimator(session, oc.w, oc.h, oc.levels)
        self.gpe.setPoints(self.pts, self.colors.T, 1.0)
        self.T_true = mat4_from_pose7(oracle, oc.case["pose7_true"])

    def close(self):
        self.gpe.close()
        self.f_new.close()


@pytest.fixture(scope="module")
def pc(session, oracle):
    c = PeCase(session, oracle, 3)
    yield c
    c.close()


@pytest.mark.parametrize("lvl", [0, 1, 2])
def test_eval_matches_oracle(pc, lvl):
    oc = pc.oc
    rng = np.random.default_rng(lvl)
    for k, (T, aff, cutoff) in enumerate(((np.eye(4), (0.0, 0.0), 20.0), (pc.T_true, (0.03, 4.0), 20.0), (pc.T_true, (-0.02, -3.0), 5.0),
                                          (pc.T_true, (0.0, 0.0), 160.0))):
        res_o, n_o, H_o, b_o, acc_o = pc.ope.calc_res(lvl, 1, T, aff, cutoff)
        res_s, n_s, H_s, b_s, _ = pc.ope.calc_res(lvl, 0, T, aff, cutoff)
        g = pc.gpe.calcResAndGS(pc.f_new, oc.K, lvl, T, aff, cutoff)
        assert g["n"] == n_o
        assert g["res6"][1] == res_o[1] and g["res6"][5] == res_o[5]
        assert rel_err(g["acc48"][:45], acc_o) < TOL_SUM
        assert abs(g["res6"][0] - res_o[0]) <= TOL_SUM * abs(res_o[0])
        assert np.allclose(g["res6"][[2, 4]], res_o[[2, 4]], rtol=1e-12, atol=0)
        assert rel_err(g["H"], H_o) < TOL_SUM and rel_err(g["b"], b_o) < TOL_SUM
        assert rel_err(g["H"], H_s) < 1e-4 and rel_err(g["b"], b_s) < 1e-3  # the reference's own fp32 summation noise


def _check_estimate(pc, T0, coarsest=None):
    oc = pc.oc
    coarsest = oc.levels - 1 if coarsest is None else coarsest
    ok_o, T_o, err_o, inl_o = pc.ope.estimate(1, T0, coarsest)
    tr_o = pc.ope.trace()
    ok_g, T_g, err_g = pc.gpe.estimate(None, 1.0, pc.f_new, oc.K, coarsest, T0)
    tr_g = pc.gpe.trace()
    assert tr_g.shape == tr_o.shape
    assert np.array_equal(tr_g[:, :4], tr_o[:, :4])  # level, iteration, accept, padded inlier count
    it = tr_o[:, 1] >= 0
    assert rel_err(tr_g[it, 7:15], tr_o[it, 7:15]) < TOL_INC
    assert np.allclose(tr_g[:, 5:7], tr_o[:, 5:7], rtol=1e-9, atol=0)
    assert ok_g == ok_o and pc.gpe.inlier_percent == inl_o
    assert abs(err_g - err_o) <= 1e-6 * abs(err_o)
    assert np.abs(T_g - T_o).max() < 1e-8
    return ok_g, T_g, err_g


def test_estimate_from_identity(pc):
    ok, T, err = _check_estimate(pc, np.eye(4))
    t_true = pc.oc.case["t_true"]
    assert np.linalg.norm(T[:3, 3] - t_true) < 0.03 * np.linalg.norm(t_true) + 2e-3


def test_estimate_from_truth_and_coarse_start(pc):
    _check_estimate(pc, pc.T_true)
    _check_estimate(pc, np.eye(4), coarsest=1)


@pytest.mark.parametrize("seed,motion", [(5, 0.5), (8, 1.5)])
def test_estimate_other_scenes(session, oracle, seed, motion):
    c = PeCase(session, oracle, seed, motion)
    try:
        _check_estimate(c, np.eye(4))
    finally:
        c.close()


def test_estimate_rejects_bad_candidate(session, oracle):
    """A wrong loop candidate (points of another scene): the alignment ends with a high residual or few inliers and
    estimate() returns false on both sides."""
    c = PeCase(session, oracle, 4)
    other = OracleCase(oracle, "tiny", 11)
    try:
        pts, colors = loop_closure_points(other, 1500, 11)
        c.ope.set_points(pts, colors, 1.0)
        c.gpe.setPoints(pts, colors.T, 1.0)
        ok_o, _, _, _ = c.ope.estimate(1, np.eye(4), c.oc.levels - 1)
        ok_g, _, _ = c.gpe.estimate(None, 1.0, c.f_new, c.oc.K, c.oc.levels - 1, np.eye(4))
        assert not ok_o and not ok_g
    finally:
        c.close()


@pytest.mark.parametrize("modeA,modeB", [(-1, -1), (0, -1), (-1, 0), (1, 1)])
def test_estimate_affine_modes(pc, modeA, modeB):
    try:
        pc.ope.set_aff_mode(modeA, modeB)
        pc.gpe.setAffineOptMode(modeA, modeB)
        _check_estimate(pc, np.eye(4))
    finally:
        pc.ope.set_aff_mode(0, 0)
        pc.gpe.setAffineOptMode(0, 0)


def test_set_points_regrow_and_exposure(session, oracle):
    """Points re-uploaded with a different count and exposure ratio (a different matched keyframe)."""
    c = PeCase(session, oracle, 6, n=700)
    try:
        _check_estimate(c, np.eye(4))
        pts, colors = loop_closure_points(c.oc, 2500, 1)
        c.ope.set_points(pts, colors, 1.3)
        c.gpe.setPoints(pts, colors.T, 1.3)
        _check_estimate(c, np.eye(4))
    finally:
        c.close()


def test_errors(session, oracle):
    from direct_stereo_slam_b200 import api

    pe = api.PoseEstimator(session, 320, 240, 3)
    f = api.FrameHessian(session, 320, 240, 3)
    f.makeImages(np.zeros((240, 320), np.float32), host=False)
    with pytest.raises(RuntimeError):
        pe.estimate(None, 1.0, f, (200, 200, 160, 120), 2, np.eye(4))  # no points yet
    with pytest.raises(RuntimeError):
        pe.setPoints(np.zeros((0, 3)), np.zeros((0, 3), np.float32), 1.0)
    pe.close()
    f.close()
