"""BASELINE.json configs[0] restated: KITTI-size stereo pairs, ScaleOptimizer only — the reference's CPU-runnable
case.  The template's inverse depths carry an unknown scale error s*; FrontEnd::optimizeScale (src/FrontEnd.cpp:975-1003) runs
TrackerAndScaler::optimizeScale from the 8 seeds {0.1, 1, 5, 10, 15, 25, 30, 50} and keeps the result with the smallest positive
RMSE.  CPU: the reference's own TrackerAndScaler.cpp compiled in place and the oracle agree bit for bit on every seed and recover
s*.  GPU: the 8 seeds in one lock step (dslam_optimize_scale_multi) give the same selection and the same scale (1e-6 relative)."""
import os

import numpy as np
import pytest

import oracle as orc
from helpers import OracleCase

SEEDS = np.array([0.1, 1, 5, 10, 15, 25, 30, 50], np.float32)
PAIRS = [(1000, 0.35), (1001, 2.7)]  # (scene seed, injected scale error s*)


def select(rmse, scale):
    """src/FrontEnd.cpp:995-1003: smallest positive error wins, first one on ties; (-1, 1.0) if none is positive."""
    best_err, best_scale = -1.0, 1.0
    for e, s in zip(rmse, scale):
        if e > 0 and (best_err < 0 or best_err > e):
            best_err, best_scale = float(e), float(s)
    return best_err, best_scale


@pytest.fixture(scope="module", params=PAIRS, ids=lambda p: "seed%d_s%.2f" % p)
def pair(request, oracle):
    seed, s_true = request.param
    return OracleCase(oracle, "kitti", seed, scale_error=s_true), s_true


def _oracle_seeds(oc, mode):
    out = [oc.trk.optimize_scale(mode, float(s0), oc.levels - 1) for s0 in SEEDS]
    return np.array([o[0] for o in out], np.float32), np.array([o[1] for o in out], np.float32)


def test_cpu_reference_and_oracle_recover_the_scale(pair):
    oc, s_true = pair
    rmse_o, scale_o = _oracle_seeds(oc, 0)  # mode 0 = the reference's fp32 / SSE accumulation order
    err, s = select(rmse_o, scale_o)
    assert err > 0 and abs(s - s_true) < 0.03 * s_true
    if not orc.ReferenceTracker.available():
        pytest.skip("oracle/_ref not built (the reference's own source is only compiled where /root/reference exists)")
    c = oc.case
    hdif = np.full(len(c["pu"]), 1e-3, np.float32)  # weight sqrt(1e-3 / (HdiF + 1e-12)) = the synthetic weights
    w = orc.ReferenceTracker.weight_from_hdif(hdif)
    oc.trk.make_coarse_depth(c["pu"], c["pv"], c["pid"], w, oc.dIp_ref)
    rt = orc.ReferenceTracker(oc.w, oc.h, oc.levels, oc.K, oc.K, oc.T_stereo)
    rt.set_ref(oc.dIp_ref, c["pu"], c["pv"], c["pid"], hdif)
    rt.set_right_frame(oc.dIp_right)
    rmse_o, scale_o = _oracle_seeds(oc, 0)
    for k, s0 in enumerate(SEEDS):
        rmse_r, scale_r = rt.optimize_scale(float(s0), oc.levels - 1)
        assert np.float32(scale_r).view(np.uint32) == scale_o[k].view(np.uint32)
        assert np.float32(rmse_r).view(np.uint32) == rmse_o[k].view(np.uint32) or (np.isnan(rmse_r) and np.isnan(rmse_o[k]))
    # restore the fixture's template (weights = the case's own)
    oc.trk.make_coarse_depth(c["pu"], c["pv"], c["pid"], c["pw"], oc.dIp_ref)


@pytest.mark.gpu
def test_gpu_seeds_in_one_lock_step(pair, session):
    from helpers import GpuCase

    oc, s_true = pair
    gc = GpuCase(session, oc, template="device")
    rmse_g, scale_g = gc.trk.optimizeScaleMulti(gc.f_right, SEEDS, oc.levels - 1)
    rmse_o, scale_o = _oracle_seeds(oc, 1)  # mode 1 = fp64 accumulation, the GPU's arithmetic
    for k in range(len(SEEDS)):
        assert abs(scale_g[k] - scale_o[k]) <= 1e-6 * abs(scale_o[k])
        assert abs(rmse_g[k] - rmse_o[k]) <= 1e-6 * abs(rmse_o[k]) or (np.isnan(rmse_g[k]) and np.isnan(rmse_o[k]))
    (err_g, s_g), (err_o, s_o) = select(rmse_g, scale_g), select(rmse_o, scale_o)
    assert int(np.argmin(np.where(rmse_g > 0, rmse_g, np.inf))) == int(np.argmin(np.where(rmse_o > 0, rmse_o, np.inf)))
    assert abs(s_g - s_o) <= 1e-6 * abs(s_o) and abs(s_g - s_true) < 0.03 * s_true
    gc.close()
