"""Fused CUDA residual / Jacobian / normal-equation kernels and the LM drivers vs the CPU oracle.

Tolerances (north star: per-iteration pose delta within 1e-5 relative of the reference):
  * integer outputs (numTermsInE, numSaturated, padded inlier count): exact;
  * raw sums against the oracle's fp64-accumulating variant (same fp32 per-point terms): 1e-11 relative — only the
    order of the fp64 additions differs;
  * against the oracle's SSE-order fp32 variant (the reference's own arithmetic): the distance is the reference's
    summation noise and is only bounded loosely (1e-4);
  * LM increments per iteration: 1e-5 relative (measured ~1e-10), identical accept/reject sequences.
"""
import numpy as np
import pytest

from helpers import IDENT7, GpuCase, OracleCase, perturbed_pose, rel_err

pytestmark = pytest.mark.gpu

TOL_SUM = 1e-11
TOL_INC = 1e-5


@pytest.fixture(scope="module")
def cases(session, oracle):
    oc = OracleCase(oracle, "tiny", 3)
    gc = GpuCase(session, oc)
    yield oc, gc
    gc.close()


def _check_pose_eval(oc, gc, lvl, pose, aff, cutoff):
    oc.trk.set_res_acc_mode(1)
    res_o, n_o = oc.trk.calc_res_pose(lvl, pose, aff, cutoff)
    H_o, b_o, acc_o = oc.trk.calc_gs_pose(lvl, 1, aff)
    oc.trk.set_res_acc_mode(0)
    res_s, _ = oc.trk.calc_res_pose(lvl, pose, aff, cutoff)
    H_s, b_s, _ = oc.trk.calc_gs_pose(lvl, 0, aff)
    g = gc.trk.calcResAndGSPose(gc.f_new, lvl, pose, aff, cutoff)
    assert g["n"][0] == n_o
    assert g["res6"][0, 1] == res_o[1]  # numTermsInE
    assert g["res6"][0, 5] == res_o[5]  # saturated ratio (float division of two exact integers)
    if n_o == 0:
        return
    assert rel_err(g["acc48"][0, :45], acc_o) < TOL_SUM
    assert abs(g["res6"][0, 0] - res_o[0]) <= TOL_SUM * abs(res_o[0])
    assert np.allclose(g["res6"][0, [2, 4]], res_o[[2, 4]], rtol=1e-12, atol=0)
    assert rel_err(g["H"][0], H_o) < TOL_SUM and rel_err(g["b"][0], b_o) < TOL_SUM
    # distance to the reference's own fp32 SSE-order accumulation = its summation noise
    assert rel_err(g["H"][0], H_s) < 1e-4 and rel_err(g["b"][0], b_s) < 1e-3
    assert abs(g["res6"][0, 0] - res_s[0]) <= 1e-4 * abs(res_s[0])


@pytest.mark.parametrize("lvl", [0, 1, 2])
def test_pose_eval_matches_oracle(cases, oracle, lvl):
    oc, gc = cases
    rng = np.random.default_rng(lvl)
    true = oc.case["pose7_true"]
    _check_pose_eval(oc, gc, lvl, IDENT7, (0.0, 0.0), 20.0)
    _check_pose_eval(oc, gc, lvl, true, (0.03, 4.0), 20.0)
    for k in range(4):
        _check_pose_eval(oc, gc, lvl, perturbed_pose(oracle, true, rng), rng.normal(0, [0.05, 5.0]), [20.0, 40.0, 5.0, 160.0][k])


def test_pose_eval_out_of_view(cases, oracle):
    """A pose that throws every point out of the image: no terms, NaN ratio, n = 0 (the reference divides by zero too)."""
    oc, gc = cases
    far = oracle.se3_exp([1.0e5, 0, 0, 0, 0, 0])
    res_o, n_o = oc.trk.calc_res_pose(0, far, (0, 0), 20.0)
    g = gc.trk.calcResAndGSPose(gc.f_new, 0, far, (0, 0), 20.0)
    assert n_o == 0 and g["n"][0] == 0 and g["res6"][0, 1] == 0
    assert np.isnan(g["res6"][0, 5]) and np.isnan(res_o[5])


def test_pose_eval_batched_equals_single(cases, oracle):
    oc, gc = cases
    rng = np.random.default_rng(0)
    nb = 150  # > one launch of 128 items
    poses = np.stack([perturbed_pose(oracle, oc.case["pose7_true"], rng) for _ in range(nb)])
    affs = rng.normal(0, [0.02, 3.0], (nb, 2))
    gb = gc.trk.calcResAndGSPose(gc.f_new, 1, poses, affs)
    for i in (0, 1, 77, 127, 128, 149):
        g1 = gc.trk.calcResAndGSPose(gc.f_new, 1, poses[i], affs[i])
        assert g1["n"][0] == gb["n"][i]
        assert rel_err(gb["acc48"][i], g1["acc48"][0]) < 1e-13
        oc.trk.set_res_acc_mode(1)
        res_o, n_o = oc.trk.calc_res_pose(1, poses[i], affs[i], 20.0)
        _, _, acc_o = oc.trk.calc_gs_pose(1, 1, affs[i])
        assert n_o == gb["n"][i] and rel_err(gb["acc48"][i, :45], acc_o) < TOL_SUM


def test_pose_eval_is_deterministic(cases):
    oc, gc = cases
    a = gc.trk.calcResAndGSPose(gc.f_new, 0, oc.case["pose7_true"], (0.01, 1.0))
    for _ in range(5):
        b = gc.trk.calcResAndGSPose(gc.f_new, 0, oc.case["pose7_true"], (0.01, 1.0))
        assert np.array_equal(a["acc48"], b["acc48"])


def _compare_traces(tg, to, ninc):
    assert tg.shape == to.shape, "different number of LM iterations: %s vs %s" % (tg.shape, to.shape)
    assert np.array_equal(tg[:, :4], to[:, :4]), "level / iteration / accept / inlier-count sequences differ"
    assert np.allclose(tg[:, 4], to[:, 4], rtol=1e-7)  # lambda (float in both)
    worst = 0.0
    for rg, ro in zip(tg, to):
        if ro[1] < 0:
            continue
        worst = max(worst, rel_err(rg[7:7 + ninc], ro[7:7 + ninc]))
    assert worst < TOL_INC, "per-iteration increment differs by %.3e relative" % worst
    assert np.allclose(tg[:, 5:7], to[:, 5:7], rtol=1e-9)
    return worst


@pytest.mark.parametrize("seed,motion", [(3, 1.0), (5, 0.5), (8, 1.5)])
def test_track_newest_coarse_matches_oracle(session, oracle, seed, motion):
    oc = OracleCase(oracle, "tiny", seed, motion_scale=motion)
    gc = GpuCase(session, oc)
    ok_o, pose_o, aff_o, last_o, flow_o = oc.trk.track_newest_coarse(1, IDENT7, (0, 0), oc.levels - 1)
    to = oc.trk.trace()
    ok_g, pose_g, aff_g, last_g = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)
    tg = gc.trk.trace()
    assert ok_g == ok_o
    worst = _compare_traces(tg, to, 8)
    assert rel_err(pose_g, pose_o) < 1e-8 and np.allclose(aff_g, aff_o, rtol=1e-7, atol=1e-9)
    assert np.allclose(last_g, last_o, rtol=1e-6, equal_nan=True)
    assert np.allclose(gc.trk.lastFlowIndicators, flow_o, rtol=1e-9)
    # against the reference's own fp32 accumulation order the LM path may differ in the last digits only
    ok_s, pose_s, aff_s, last_s, _ = oc.trk.track_newest_coarse(0, IDENT7, (0, 0), oc.levels - 1)
    assert ok_s == ok_g and rel_err(pose_g, pose_s) < 1e-4
    print("worst per-iteration increment rel. error: %.2e; pose vs SSE-order oracle: %.2e" % (worst, rel_err(pose_g, pose_s)))
    gc.close()


def test_track_abort_leaves_outputs_untouched(cases):
    oc, gc = cases
    start = np.array(IDENT7)
    ok_o, pose_o, aff_o, last_o, _ = oc.trk.track_newest_coarse(1, start, (0, 0), oc.levels - 1, min_res=np.full(5, 0.1))
    ok_g, pose_g, aff_g, last_g = gc.trk.trackNewestCoarse(gc.f_new, start, (0, 0), oc.levels - 1, minResForAbort=np.full(5, 0.1))
    assert not ok_o and not ok_g
    assert np.array_equal(pose_g, start) and np.array_equal(aff_g, [0, 0])
    assert np.allclose(last_g, last_o, rtol=1e-6, equal_nan=True)


@pytest.mark.parametrize("modeA,modeB", [(-1, -1), (0, -1), (-1, 0), (1, 1)])
def test_track_affine_modes(session, oracle, modeA, modeB):
    oc = OracleCase(oracle, "tiny", 6)
    gc = GpuCase(session, oc)
    oc.trk.set_aff_mode(modeA, modeB)
    gc.trk.setAffineOptMode(modeA, modeB)
    ok_o, pose_o, aff_o, last_o, _ = oc.trk.track_newest_coarse(1, IDENT7, (0, 0), oc.levels - 1)
    ok_g, pose_g, aff_g, last_g = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)
    assert ok_g == ok_o
    _compare_traces(gc.trk.trace(), oc.trk.trace(), 8)
    assert rel_err(pose_g, pose_o) < 1e-8 and np.allclose(aff_g, aff_o, rtol=1e-7, atol=1e-9)
    gc.close()


def test_track_multi_equals_sequential(cases, oracle):
    """Lock-step hypotheses (one launch per LM round for all of them) give exactly the sequential results."""
    oc, gc = cases
    rng = np.random.default_rng(4)
    starts = np.stack([IDENT7] + [perturbed_pose(oracle, IDENT7, rng, 0.05, 0.01) for _ in range(6)])
    affs = np.zeros((7, 2))
    ok_m, pose_m, aff_m, last_m, flow_m = gc.trk.trackNewestCoarseMulti(gc.f_new, starts, affs, oc.levels - 1)
    for i in range(7):
        ok_1, pose_1, aff_1, last_1 = gc.trk.trackNewestCoarse(gc.f_new, starts[i], affs[i], oc.levels - 1)
        assert ok_1 == ok_m[i]
        assert rel_err(pose_m[i], pose_1) < 1e-9 and np.allclose(aff_m[i], aff_1, rtol=1e-8, atol=1e-10)
        assert np.allclose(last_m[i], last_1, rtol=1e-7, equal_nan=True)


# ---- scale ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lvl", [0, 1, 2])
def test_scale_eval_matches_oracle(cases, lvl):
    oc, gc = cases
    for scale, cutoff in [(1.0, 20.0), (0.5, 20.0), (2.0, 40.0), (10.0, 160.0), (0.1, 20.0)]:
        oc.trk.set_res_acc_mode(1)
        res_o, n_o = oc.trk.calc_res_scale(lvl, scale, cutoff)
        H_o, b_o, acc_o = oc.trk.calc_gs_scale(lvl, 1, scale)
        H_s, b_s, _ = oc.trk.calc_gs_scale(lvl, 0, scale)
        g = gc.trk.calcResAndGSScale(gc.f_right, lvl, scale, cutoff)
        assert g["n"][0] == n_o and g["res6"][0, 1] == res_o[1] and g["res6"][0, 5] == res_o[5]
        assert rel_err(g["acc8"][0, :3], acc_o) < TOL_SUM
        assert abs(g["res6"][0, 0] - res_o[0]) <= TOL_SUM * abs(res_o[0])
        assert g["H"][0] == np.float32(H_o) and g["b"][0] == np.float32(b_o) or (
            abs(g["H"][0] - H_o) <= 2e-7 * abs(H_o) and abs(g["b"][0] - b_o) <= 2e-7 * abs(b_o) + 1e-12)
        assert abs(g["H"][0] - H_s) <= 1e-4 * abs(H_s)


@pytest.mark.parametrize("seed,scale_error,seed_scale", [(4, 2.5, 1.0), (4, 2.5, 5.0), (7, 0.4, 1.0), (9, 1.0, 0.1), (9, 1.0, 30.0)])
def test_optimize_scale_matches_oracle(session, oracle, seed, scale_error, seed_scale):
    oc = OracleCase(oracle, "tiny", seed, scale_error=scale_error)
    gc = GpuCase(session, oc)
    rmse_o, s_o = oc.trk.optimize_scale(1, seed_scale, oc.levels - 1)
    to = oc.trk.trace()
    rmse_g, s_g = gc.trk.optimizeScale(gc.f_right, seed_scale, oc.levels - 1)
    tg = gc.trk.trace()
    assert tg.shape == to.shape and np.array_equal(tg[:, :4], to[:, :4])
    for rg, ro in zip(tg, to):
        if ro[1] >= 0 and ro[7] != 0:
            assert abs(rg[7] - ro[7]) <= TOL_INC * abs(ro[7])
    assert abs(s_g - s_o) <= 1e-6 * abs(s_o) and (abs(rmse_g - rmse_o) <= 1e-6 * abs(rmse_o) or (np.isnan(rmse_g) and np.isnan(rmse_o)))
    gc.close()


def test_optimize_scale_multi_equals_sequential(session, oracle):
    """The 8 seeds of FrontEnd::optimizeScale (src/FrontEnd.cpp:995-1003) in lock step."""
    oc = OracleCase(oracle, "tiny", 4, scale_error=2.5)
    gc = GpuCase(session, oc)
    seeds = np.array([0.1, 1, 5, 10, 15, 25, 30, 50], np.float32)
    rmse_m, s_m = gc.trk.optimizeScaleMulti(gc.f_right, seeds, oc.levels - 1)
    for i, s0 in enumerate(seeds):
        rmse_1, s_1 = gc.trk.optimizeScale(gc.f_right, float(s0), oc.levels - 1)
        assert s_1 == s_m[i] and (rmse_1 == rmse_m[i] or (np.isnan(rmse_1) and np.isnan(rmse_m[i])))
        rmse_o, s_o = oc.trk.optimize_scale(1, float(s0), oc.levels - 1)
        assert abs(s_1 - s_o) <= 1e-6 * abs(s_o)
    gc.close()


def test_scale_idepth_and_template_roundtrip(session, oracle):
    oc = OracleCase(oracle, "tiny", 4, scale_error=2.5)
    gc = GpuCase(session, oc)
    gc.trk.scaleCoarseDepthL0(2.5)
    oc.trk.scale_idepth(2.5)
    for l in range(oc.levels):
        g = gc.trk.ref_level(l)
        o = oc.trk.get_ref_level(l)
        for a, b in zip(g, o):
            assert np.array_equal(a, b)
    rmse_o, s_o = oc.trk.optimize_scale(1, 1.0, oc.levels - 1)
    rmse_g, s_g = gc.trk.optimizeScale(gc.f_right, 1.0, oc.levels - 1)
    assert abs(s_g - s_o) <= 1e-6 * abs(s_o) and abs(s_g - 1.0) < 0.05
    gc.close()


def test_template_built_on_device(session, oracle):
    """dslam_ref_build (makeCoarseDepthL0 on the GPU) reproduces the host loop's pc_* arrays in the same order."""
    for cfg, seed in (("tiny", 3), ("tiny", 12)):
        oc = OracleCase(oracle, cfg, seed)
        gc = GpuCase(session, oc, template="device")
        for l in range(oc.levels):
            g = gc.trk.ref_level(l)
            o = oc.ref_levels[l]
            assert len(g[0]) == len(o[0]) == gc.pc_n[l]
            for a, b in zip(g, o):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        ok_o, pose_o, aff_o, _, _ = oc.trk.track_newest_coarse(1, IDENT7, (0, 0), oc.levels - 1)
        ok_g, pose_g, aff_g, _ = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)
        assert ok_g == ok_o and rel_err(pose_g, pose_o) < 1e-8
        gc.close()


def test_errors(session, oracle):
    from direct_stereo_slam_b200 import api
    from direct_stereo_slam_b200._lib import ESTATE, EINVAL, DslamError

    trk = api.TrackerAndScaler(session, 320, 192, np.eye(4).reshape(-1), [200, 200, 160, 96], levels=3)
    fr = api.FrameHessian(session, 320, 192, 3)
    with pytest.raises(DslamError) as e:
        trk.trackNewestCoarse(fr, IDENT7, (0, 0), 2)
    assert e.value.code == ESTATE  # no reference yet
    trk.setCoarseTrackingRefArrays([[np.zeros(4, np.float32)] * 4] * 3)
    with pytest.raises(DslamError) as e:
        trk.trackNewestCoarse(fr, IDENT7, (0, 0), 2)
    assert e.value.code == ESTATE  # frame not built
    fr.makeImages(np.zeros((192, 320), np.float32), host=False)
    with pytest.raises(DslamError) as e:
        trk.trackNewestCoarse(fr, IDENT7, (0, 0), 3)
    assert e.value.code == EINVAL  # coarsestLvl out of range
    trk.close()
    fr.close()


@pytest.mark.parametrize("scenario", ["first_try_wins", "needs_retries", "nothing_good", "improves_later"])
def test_track_new_coarse_speculative_equals_sequential(session, oracle, scenario):
    """dslam_track_new_coarse (the hypothesis loop of FrontEnd::trackNewCoarse, src/FrontEnd.cpp:192-252, evaluated
    speculatively in lock-step batches) returns what the sequential loop returns."""
    from helpers import rotation_hypotheses

    oc = OracleCase(oracle, "tiny", 3)
    gc = GpuCase(session, oc)
    true = oc.case["pose7_true"]
    off = oracle.se3_mul(oracle.se3_exp([0.0, 0, 0, 0.0, 0.09, 0.0]), true)  # 5 degrees off: fails on the fine levels
    if scenario == "first_try_wins":
        tries = np.concatenate([[true, IDENT7], rotation_hypotheses(oracle, true)])
        last = np.full(5, 20.0)
    elif scenario == "needs_retries":
        tries = np.concatenate([[off, oracle.se3_mul(off, off)], rotation_hypotheses(oracle, off, (0.02, 0.03, 0.045)), [true]])
        last = np.full(5, 7.0)
    elif scenario == "nothing_good":
        far = oracle.se3_exp([1.0e5, 0, 0, 0, 0, 0])
        tries = np.stack([far, oracle.se3_mul(far, far), far])
        last = np.full(5, 1.0)
    else:
        tries = np.concatenate([[IDENT7, off], rotation_hypotheses(oracle, off, (0.045,)), [true, true]])
        last = np.full(5, 0.5)  # never good enough: every hypothesis is tried
    ref = oc.trk.track_new_coarse(1, tries, (0.0, 0.0), oc.levels - 1, last)
    l0 = session.launch_count()
    got = gc.trk.trackNewCoarse(gc.f_new, tries, (0.0, 0.0), oc.levels - 1, last)
    launches = session.launch_count() - l0
    assert got["haveOneGood"] == ref["haveOneGood"] and got["tryIterations"] == ref["tryIterations"], (got, ref)
    assert rel_err(got["pose"], ref["pose"]) < 1e-8 and np.allclose(got["aff"], ref["aff"], rtol=1e-7, atol=1e-9)
    assert np.allclose(got["achievedRes"], ref["achievedRes"], rtol=1e-6, equal_nan=True)
    assert np.allclose(got["flow"], ref["flow"], rtol=1e-8)
    if scenario == "first_try_wins":
        assert got["tryIterations"] == 1 and launches < 80  # only the first hypothesis was evaluated
    if scenario == "improves_later":
        assert got["tryIterations"] == len(tries)
    print(scenario, "tries", got["tryIterations"], "of", len(tries), "launches", launches)
    gc.close()
