"""Parity at BASELINE.json's full sizes (configs 1-3): KITTI 1232x368 / 2000 active points, Malaga 1024x768, synthetic
1920x1200 / 8000 active points, 5 pyramid levels each — the whole chain (device-built template, fused evaluation at every
level, the complete trackNewestCoarse and optimizeScale LM loops, the batched lock step) against the CPU oracle on the same
seeded inputs.  The oracle needs ~10-60 ms per frame at these sizes, so this is direct parity, not only properties.

Tolerances as in test_gpu_tracker.py: integers exact, fp64 sums 1e-11 relative, LM increments 1e-5 relative with identical
accept / reject sequences, scale 1e-6 relative."""
import numpy as np
import pytest

from direct_stereo_slam_b200 import api
from helpers import IDENT7, GpuCase, OracleCase, perturbed_pose, rel_err
from test_gpu_tracker import TOL_INC, TOL_SUM, _check_pose_eval, _compare_traces

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["kitti", "malaga", "synth1920"])
def full(request, session, oracle):
    oc = OracleCase(oracle, request.param, 3, scale_error=1.6)
    gc = GpuCase(session, oc, template="device")
    yield oc, gc
    gc.close()


def test_device_template_equals_oracle(full):
    """makeCoarseDepthL0 on the device (dslam_ref_build): same point count, order and records at every level."""
    oc, gc = full
    assert oc.levels == 5
    for l in range(oc.levels):
        got = gc.trk.ref_level(l)
        want = oc.ref_levels[l]
        assert got[0].size == want[0].size
        for a, b in zip(got, want):
            assert np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("lvl", [0, 2, 4])
def test_pose_eval_full_size(full, oracle, lvl):
    oc, gc = full
    rng = np.random.default_rng(10 + lvl)
    true = oc.case["pose7_true"]
    _check_pose_eval(oc, gc, lvl, IDENT7, (0.0, 0.0), 20.0)
    _check_pose_eval(oc, gc, lvl, true, (0.03, 4.0), 20.0)
    _check_pose_eval(oc, gc, lvl, perturbed_pose(oracle, true, rng), rng.normal(0, [0.05, 5.0]), 40.0)


def test_track_and_scale_full_size(full):
    oc, gc = full
    ok_o, pose_o, aff_o, last_o, flow_o = oc.trk.track_newest_coarse(1, IDENT7, (0, 0), oc.levels - 1)
    to = oc.trk.trace()
    ok_g, pose_g, aff_g, last_g = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)
    tg = gc.trk.trace()
    assert ok_g == ok_o and ok_o
    worst = _compare_traces(tg, to, 8)
    assert worst < TOL_INC
    assert rel_err(pose_g, pose_o) < 1e-8 and np.allclose(aff_g, aff_o, rtol=1e-7, atol=1e-9)
    assert np.allclose(last_g, last_o, rtol=1e-6, equal_nan=True)
    assert np.allclose(gc.trk.lastFlowIndicators, flow_o, rtol=1e-9)
    # the tracker recovers the ground-truth motion of the synthetic scene (size-independent property); the template's inverse
    # depths carry the injected scale error, so the monocular translation comes out divided by it
    assert rel_err(pose_g[4:] * oc.case["scale_error"], oc.case["pose7_true"][4:]) < 0.05
    assert rel_err(pose_g[:4], oc.case["pose7_true"][:4]) < 1e-3
    rmse_o, s_o = oc.trk.optimize_scale(1, 1.0, oc.levels - 1)
    to = oc.trk.trace()
    rmse_g, s_g = gc.trk.optimizeScale(gc.f_right, 1.0, oc.levels - 1)
    tg = gc.trk.trace()
    assert tg.shape == to.shape and np.array_equal(tg[:, :4], to[:, :4])
    assert abs(s_g - s_o) <= 1e-6 * abs(s_o) and abs(rmse_g - rmse_o) <= 1e-6 * abs(rmse_o)
    assert abs(s_g - oc.case["scale_error"]) < 0.05 * oc.case["scale_error"]  # ... and the injected scale error


def test_lock_step_batch_full_size(full):
    """Tracking and scale optimisation of the frame in ONE lock step, with the pyramid built on the overlap stream: same
    results as the separate calls."""
    oc, gc = full
    ok1, pose1, aff1, last1 = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)
    rmse1, s1 = gc.trk.optimizeScale(gc.f_right, 1.0, oc.levels - 1)
    f2 = api.FrameHessian(gc.trk.s, oc.w, oc.h, oc.levels)
    f2.upload(oc.case["img_new"])
    api.build_frames([f2], overlap=True)
    ok, poses, affs, last, rmse, scales = api.lm_batch([gc.trk], [f2], IDENT7[None], np.zeros((1, 2)), oc.levels - 1, [gc.trk], [gc.f_right], [1.0])
    assert ok[0] == ok1 and rel_err(poses[0], pose1) < 1e-9 and np.allclose(affs[0], aff1, rtol=1e-8, atol=1e-10)
    assert scales[0] == np.float32(s1) and rmse[0] == np.float32(rmse1)
    f2.close()
