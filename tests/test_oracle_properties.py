"""Independent checks of the oracle's building blocks (numpy / scipy restatements, analytic properties)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle as orc
from helpers import IDENT7, OracleCase, rel_err
from direct_stereo_slam_b200 import synthetic as syn


def test_se3_exp_and_product(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        xi = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.7, 3)])
        p = oracle.se3_exp(xi)
        R, t = syn.se3_exp_mat(xi)
        assert np.allclose(oracle.se3_R(p), R, atol=1e-12) and np.allclose(p[4:], t, atol=1e-12)
        assert np.allclose(Rotation.from_quat(p[:4]).as_matrix(), R, atol=1e-12)
        q = oracle.se3_exp(rng.normal(0, 0.3, 6))
        pq = oracle.se3_mul(p, q)
        assert np.allclose(oracle.se3_R(pq), R @ oracle.se3_R(q), atol=1e-12)
        assert np.allclose(pq[4:], R @ q[4:] + t, atol=1e-12)
    tiny = oracle.se3_exp([1e-3, 2e-3, 3e-3, 1e-12, 0, 0])  # small-angle branch (theta < 1e-10)
    assert np.allclose(tiny[4:], [1e-3, 2e-3, 3e-3], atol=1e-14) and abs(tiny[3] - 1) < 1e-15


def test_ldlt_solve(oracle):
    rng = np.random.default_rng(1)
    for n in (6, 7, 8):
        for _ in range(20):
            A = rng.normal(size=(n, n))
            A = A @ A.T + 1e-3 * np.eye(n)
            b = rng.normal(size=n)
            assert np.allclose(oracle.ldlt_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)


def test_make_images_against_numpy(oracle):
    rng = np.random.default_rng(2)
    w, h, levels = 96, 64, 3
    img = rng.uniform(0, 255, (h, w)).astype(np.float32)
    dIp, ag = oracle.make_images(img, levels)
    off = orc.level_offsets(w, h, levels)
    I = img
    for l in range(levels):
        wl, hl = w >> l, h >> l
        if l > 0:
            I = (np.float32(0.25) * (((I[0::2, 0::2] + I[0::2, 1::2]) + I[1::2, 0::2]) + I[1::2, 1::2])).astype(np.float32)
        d = dIp[off[l]:off[l + 1]].reshape(hl, wl, 3)
        assert np.array_equal(d[..., 0], I)
        flat = I.reshape(-1)
        idx = np.arange(wl, wl * (hl - 1))
        dx = np.float32(0.5) * (flat[idx + 1] - flat[idx - 1])  # linear index: wraps at the row ends like the reference
        dy = np.float32(0.5) * (flat[idx + wl] - flat[idx - wl])
        assert np.array_equal(d.reshape(-1, 3)[idx, 1], dx) and np.array_equal(d.reshape(-1, 3)[idx, 2], dy)
        assert np.array_equal(ag[off[l]:off[l + 1]][idx], dx * dx + dy * dy)


def test_pyr_levels_rule():
    assert orc.pyr_levels_used(1232, 368) == 5      # KITTI after the calibration crop
    assert orc.pyr_levels_used(1241, 376) == 1      # the raw width is odd
    assert orc.pyr_levels_used(1024, 768) == 6 or orc.pyr_levels_used(1024, 768) >= 5
    assert orc.pyr_levels_used(1920, 1200) == 5     # 75 is odd


def test_tracking_converges_and_modes_agree(oracle):
    oc = OracleCase(oracle, "tiny", 3)
    res = {}
    for mode in (0, 1):
        ok, pose, aff, last, flow = oc.trk.track_newest_coarse(mode, IDENT7, (0, 0), oc.levels - 1)
        assert ok
        res[mode] = pose
        t_true = oc.case["pose7_true"][4:]
        assert np.linalg.norm(pose[4:] - t_true) < 0.03 * np.linalg.norm(t_true) + 2e-3
        assert np.linalg.norm(pose[:3] - oc.case["pose7_true"][:3]) < 3e-4
    # SSE-order fp32 accumulation vs fp64 accumulation: the reference's own summation noise floor
    assert rel_err(res[0], res[1]) < 1e-5


@pytest.mark.parametrize("scale_error", [0.4, 1.0, 2.5])
def test_scale_optimizer_recovers_scale(oracle, scale_error):
    oc = OracleCase(oracle, "tiny", 4, scale_error=scale_error)
    rmse, s = oc.trk.optimize_scale(0, 1.0, oc.levels - 1)
    assert abs(s - scale_error) < 0.05 * scale_error and 0 < rmse < 15.0


def test_template_dilation_and_counts(oracle):
    oc = OracleCase(oracle, "tiny", 3)
    n0 = len(oc.ref_levels[0][0])
    npts = len(oc.case["pu"])
    assert npts < n0 <= 5 * npts  # every point dilates to at most its 4 diagonal neighbours on level 0
    for l, (u, v, idp, col) in enumerate(oc.ref_levels):
        wl, hl = oc.w >> l, oc.h >> l
        assert u.min() >= 2 and u.max() < wl - 2 and v.min() >= 2 and v.max() < hl - 2
        assert np.all(idp > 0) and np.all(np.isfinite(col))
        order = v.astype(np.int64) * wl + u.astype(np.int64)
        assert np.all(np.diff(order) > 0)  # raster order, no duplicates


def test_sc_generate_properties(oracle):
    rng = np.random.default_rng(3)
    pts = rng.normal(0, 12, (4000, 3))
    rk, si, sv, tfm = oracle.sc_generate(pts)
    assert np.all(np.diff(si) > 0) and si.max() < 1200
    dense = np.zeros(1200)
    dense[si] = sv
    norms = np.sqrt((dense.reshape(60, 20) ** 2).sum(1))
    assert np.allclose(norms[norms > 0], 1.0, atol=1e-12)  # every occupied sector column is L2-normalised
    occ = (dense.reshape(60, 20) != 0).sum(0) / 60.0
    assert np.allclose(rk, occ.astype(np.float32))
    R = tfm[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-9)
    # self distance is 0 and the search prefers the identical descriptor
    ptr = np.array([0, len(si)], np.int32)
    i, d = oracle.search_sc(si, sv, ptr, si, sv, np.array([0], np.int32))
    assert i == 0 and abs(d) < 1e-6
