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


# The north star asks for "per-iteration pose delta within 1e-5 rel of the reference".  The reference's own arithmetic (oracle
# mode 0: 4 SSE lanes x 3 tiers of fp32 sums, bit-identical to the compiled TrackerAndScaler.cpp) and the fp64-accumulating
# arithmetic the GPU shares with oracle mode 1 (the sanctioned target, BASELINE.md 4: agreement 1e-5, measured ~1e-10) differ
# by the reference's own fp32 summation noise.  That floor is asserted here — not only printed — so that a regression in
# either direction shows up: increments of a few 1e-3 relative on the near-converged iterations, poses 1e-5.
NOISE_INC_WORST = 2e-2   # worst per-iteration increment, relative, where the accept / reject sequences coincide
NOISE_INC_MEDIAN = 2e-3
NOISE_POSE = 1e-5        # final pose, relative


def test_distance_to_the_reference_faithful_trace_is_asserted(full):
    oc, gc = full
    init = perturbed_pose(oc.o, oc.case["pose7_true"], np.random.default_rng(1), trans=0.02, rot=0.004)
    ok_s, pose_s, aff_s, last_s, _ = oc.trk.track_newest_coarse(0, init, (0, 0), oc.levels - 1)  # mode 0 = the reference's arithmetic
    ts = oc.trk.trace()
    ok_g, pose_g, aff_g, last_g = gc.trk.trackNewestCoarse(gc.f_new, init, (0, 0), oc.levels - 1)
    tg = gc.trk.trace()
    assert ok_g == ok_s
    errs, same = [], 0
    for rg, rs in zip(tg, ts):
        if not np.array_equal(rg[:3], rs[:3]):  # level / iteration / accept: from the first difference on the paths are different problems
            break
        same += 1
        if rs[1] >= 0:
            errs.append(rel_err(rg[7:15], rs[7:15]))
    assert same >= 0.8 * len(ts), "accept / reject sequences diverge early: %d of %d rows" % (same, len(ts))
    assert max(errs) < NOISE_INC_WORST and np.median(errs) < NOISE_INC_MEDIAN, (max(errs), np.median(errs))
    assert rel_err(pose_g, pose_s) < NOISE_POSE and np.allclose(aff_g, aff_s, rtol=1e-4, atol=1e-4)
    assert np.allclose(last_g, last_s, rtol=1e-4, equal_nan=True)
    print("noise floor vs the reference-faithful trace: rows %d/%d, increments worst %.2e median %.2e, pose %.2e" % (
        same, len(ts), max(errs), float(np.median(errs)), rel_err(pose_g, pose_s)))


def test_hypothesis_loop_83_tries_full_size(full):
    """f-4 at full size with the 83 initialisations of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:147-180): (a) the normal case —
    the constant-motion guess wins and nothing else is evaluated; (b) a bad motion model — wrong guesses first, so the loop walks
    the rotation retries; both equal the sequential loop on the oracle."""
    from direct_stereo_slam_b200 import synthetic as syn

    oc, gc = full
    c = oc.case
    good = syn.pose7(*syn.se3_exp_mat(c["xi_true"] * 0.9))
    dbl = syn.pose7(*syn.se3_exp_mat(c["xi_true"] * 1.8))
    half = syn.pose7(*syn.se3_exp_mat(c["xi_true"] * 0.45))
    tries = syn.frontend_pose_tries(good, dbl, half, IDENT7)
    assert len(tries) == 83
    last = np.full(5, 100.0)
    ref = oc.trk.track_new_coarse(1, tries, (0.0, 0.0), oc.levels - 1, last)
    got = gc.trk.trackNewCoarse(gc.f_new, tries, (0.0, 0.0), oc.levels - 1, last)
    assert got["tryIterations"] == ref["tryIterations"] == 1 and got["haveOneGood"]
    assert rel_err(got["pose"], ref["pose"]) < 1e-8 and np.allclose(got["achievedRes"], ref["achievedRes"], rtol=1e-6, equal_nan=True)
    # (b) a threshold no try can meet (last_coarse_rmse far below what the scene allows): the loop never breaks, all 83
    # hypotheses are evaluated (the speculative stages 1, 4, "the rest") and the replay applies the level-abort test of every
    # try against the evolving achievedRes exactly like the sequential loop
    last_b = np.full(5, ref["achievedRes"][0] / 1.6)
    ref_b = oc.trk.track_new_coarse(1, tries, (0.0, 0.0), oc.levels - 1, last_b)
    got_b = gc.trk.trackNewCoarse(gc.f_new, tries, (0.0, 0.0), oc.levels - 1, last_b)
    assert got_b["tryIterations"] == ref_b["tryIterations"] == 83 and got_b["haveOneGood"] == ref_b["haveOneGood"], (got_b, ref_b)
    assert rel_err(got_b["pose"], ref_b["pose"]) < 1e-8 and np.allclose(got_b["aff"], ref_b["aff"], rtol=1e-7, atol=1e-9)
    assert np.allclose(got_b["achievedRes"], ref_b["achievedRes"], rtol=1e-6, equal_nan=True)
    assert np.allclose(got_b["flow"], ref_b["flow"], rtol=1e-8)
    # (c) a wrong motion model first (the constant-motion slot 12 degrees off), the usable guess only in the zero-motion slot
    bad = oc.o.se3_mul(oc.o.se3_exp([0.0, 0, 0, 0.0, 0.21, 0.0]), good)
    tries_c = syn.frontend_pose_tries(bad, oc.o.se3_mul(bad, bad), bad, good)
    last_c = np.full(5, ref["achievedRes"][0] / 1.4)
    ref_c = oc.trk.track_new_coarse(1, tries_c, (0.0, 0.0), oc.levels - 1, last_c)
    got_c = gc.trk.trackNewCoarse(gc.f_new, tries_c, (0.0, 0.0), oc.levels - 1, last_c)
    assert got_c["tryIterations"] == ref_c["tryIterations"] > 1 and got_c["haveOneGood"] == ref_c["haveOneGood"], (got_c, ref_c)
    assert rel_err(got_c["pose"], ref_c["pose"]) < 1e-8 and np.allclose(got_c["achievedRes"], ref_c["achievedRes"], rtol=1e-6, equal_nan=True)
    print("83-try loop: normal case 1 try; unreachable threshold %d tries; wrong motion model %d tries" % (got_b["tryIterations"], got_c["tryIterations"]))
