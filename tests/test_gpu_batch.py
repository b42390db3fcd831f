"""Batched entry points (the B200-first levers): all pyramids of a step in two launches, pose trackers and scale
optimisers of many streams in one lock step.  Each must reproduce the one-at-a-time calls."""
import numpy as np
import pytest

from helpers import IDENT7, GpuCase, OracleCase, rel_err
from direct_stereo_slam_b200 import api

pytestmark = pytest.mark.gpu


def test_build_frames_batch_equals_single(session):
    rng = np.random.default_rng(0)
    w, h, levels = 320, 192, 3
    imgs = [np.clip(rng.normal(128, 40, (h, w)), 0, 255).astype(np.float32) for _ in range(5)]
    single = []
    for im in imgs:
        f = api.FrameHessian(session, w, h, levels)
        f.makeImages(im)
        single.append((f.dIp_all.copy(), f.absSquaredGrad_all.copy()))
        f.close()
    frames = [api.FrameHessian(session, w, h, levels) for _ in imgs]
    for f, im in zip(frames, imgs):
        f.upload(im)
    api.build_frames(frames, stage_host=3)
    for f, (d, a) in zip(frames, single):
        f.download()
        assert np.array_equal(f.dIp_all.view(np.uint32), d.view(np.uint32)) and np.array_equal(f.absSquaredGrad_all.view(np.uint32), a.view(np.uint32))
    # partial mirror: only level 0 of dIp
    frames[0].dIp_all[:] = -1
    frames[0].absSquaredGrad_all[:] = -1
    frames[0].download(levels=[0], abs_grad=False)
    n0 = w * h
    assert np.array_equal(frames[0].dIp_all[:n0].view(np.uint32), single[0][0][:n0].view(np.uint32))
    assert np.all(frames[0].dIp_all[n0:] == -1) and np.all(frames[0].absSquaredGrad_all == -1)
    # mixed geometries fall into separate launches
    g = api.FrameHessian(session, 160, 96, 2)
    g.upload(imgs[0][:96, :160])
    api.build_frames([frames[1], g, frames[2]])
    g.download()
    g2 = api.FrameHessian(session, 160, 96, 2)
    g2.makeImages(imgs[0][:96, :160])
    assert np.array_equal(g.dIp_all.view(np.uint32), g2.dIp_all.view(np.uint32))
    for f in frames + [g, g2]:
        f.close()


def test_lm_batch_equals_separate_calls(session, oracle):
    """3 streams: all track, two of them also optimise the scale in the same launches."""
    ocs = [OracleCase(oracle, "tiny", sd, scale_error=se) for sd, se in ((3, 1.0), (4, 2.5), (9, 0.6))]
    gcs = [GpuCase(session, oc) for oc in ocs]
    trk = [g.trk for g in gcs]
    left = [g.f_new for g in gcs]
    sep = []
    for g, oc in zip(gcs, ocs):
        ok, pose, aff, last = g.trk.trackNewestCoarse(g.f_new, IDENT7, (0, 0), oc.levels - 1)
        rmse, s = g.trk.optimizeScale(g.f_right, 1.0, oc.levels - 1)
        sep.append((ok, pose, aff, last, rmse, s))
    before = session.launch_count()
    ok, poses, affs, last, rmse, scales = api.lm_batch(trk, left, np.tile(IDENT7, (3, 1)), np.zeros((3, 2)), ocs[0].levels - 1,
                                                       [trk[1], trk[2]], [gcs[1].f_right, gcs[2].f_right], [1.0, 1.0])
    launches = session.launch_count() - before
    for i in range(3):
        assert ok[i] == sep[i][0] and rel_err(poses[i], sep[i][1]) < 1e-9 and np.allclose(affs[i], sep[i][2], rtol=1e-8, atol=1e-10)
        assert np.allclose(last[i], sep[i][3], rtol=1e-7, equal_nan=True)
    assert scales[0] == np.float32(sep[1][5]) and scales[1] == np.float32(sep[2][5])
    assert rmse[0] == np.float32(sep[1][4]) and rmse[1] == np.float32(sep[2][4])
    # one launch per round: no more launches than the longest machine needs evaluations
    worst = max(len(g.trk.trace()) for g in gcs) + 60
    assert launches <= worst
    # the reusable plan objects (pointer arrays built once) are the same call
    plan = api.LmBatchPlan(trk, left, [trk[1], trk[2]], [gcs[1].f_right, gcs[2].f_right], ocs[0].levels - 1)
    for _ in range(2):
        ok2, poses2, affs2, last2, rmse2, scales2 = plan.run(np.tile(IDENT7, (3, 1)), np.zeros((3, 2)), [1.0, 1.0])
        assert np.array_equal(ok2, ok) and np.array_equal(poses2, poses) and np.array_equal(affs2, affs)
        assert np.array_equal(rmse2, rmse) and np.array_equal(scales2, scales)
    # and against the oracle
    for i, oc in enumerate(ocs):
        ok_o, pose_o, aff_o, _, _ = oc.trk.track_newest_coarse(1, IDENT7, (0, 0), oc.levels - 1)
        assert ok_o == ok[i] and rel_err(poses[i], pose_o) < 1e-8
    for k, i in enumerate((1, 2)):
        rmse_o, s_o = ocs[i].trk.optimize_scale(1, 1.0, ocs[i].levels - 1)
        assert abs(scales[k] - s_o) <= 1e-6 * abs(s_o)
    for g in gcs:
        g.close()


def test_resident_server_equals_launch_per_round(oracle, monkeypatch):
    """The resident evaluation server (doorbells instead of a kernel launch per LM round) and the launch-per-round path run the
    same eval_cta code on the same work split: identical traces, poses, scales — bit for bit — for one stream and for a batch
    of streams with pose and scale machines in one lock step."""
    from helpers import GpuCase, OracleCase

    results = {}
    oc = OracleCase(oracle, "tiny", 3, scale_error=1.4)
    for mode in ("1", "0"):
        monkeypatch.setenv("DSLAM_LM_SERVER", mode)
        s = api.Session(0)
        gc = GpuCase(s, oc, template="device")
        l0 = s.launch_count()
        ok, pose, aff, last = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)
        single_launches = s.launch_count() - l0
        tr = gc.trk.trace()
        rmse, sc = gc.trk.optimizeScale(gc.f_right, 1.0, oc.levels - 1)
        n = 24
        poses = np.tile(IDENT7, (n, 1))
        poses[:, 4:] += np.random.default_rng(0).normal(0, 0.01, (n, 3))
        out = api.lm_batch([gc.trk] * n, [gc.f_new] * n, poses, np.zeros((n, 2)), oc.levels - 1, [gc.trk] * 3, [gc.f_right] * 3, [0.8, 1.0, 1.3])
        results[mode] = (ok, pose, aff, last, tr, rmse, sc, out, single_launches)
        gc.close()
        s.close()
    a, b = results["1"], results["0"]
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3], equal_nan=True)
    assert np.array_equal(a[4], b[4]) and a[5] == b[5] and a[6] == b[6]
    for x, y in zip(a[7], b[7]):
        assert np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True)
    assert a[8] == 1 and b[8] > 10, "server: one launch per LM call; launch path: one per round (%d / %d)" % (a[8], b[8])


def test_speculation_over_rejection_runs_saves_round_trips(oracle, monkeypatch):
    """After a rejected LM step the next candidate depends only on (H, b, lambda): the driver evaluates the candidates of the next
    rejections of a run in the same round (DSLAM_LM_SPEC, default depth 4).  The machine still consumes exactly the sequential
    sequence of evaluations — the trace equals the oracle's and the one without speculation — in fewer host round trips."""
    from helpers import GpuCase, OracleCase
    from test_gpu_tracker import _compare_traces

    oc = OracleCase(oracle, "kitti", 1000, scale_error=1.1)
    from direct_stereo_slam_b200 import synthetic as syn
    init = syn.pose7(*syn.se3_exp_mat(oc.case["xi_true"] * 0.9))
    ok_o, pose_o, aff_o, last_o, _ = oc.trk.track_newest_coarse(1, init, (0, 0), oc.levels - 1)
    to = oc.trk.trace()
    assert int(((to[:, 1] >= 0) & (to[:, 2] == 0)).sum()) >= 4, "the scene is expected to produce a run of rejected steps"
    rmse_o, s_o = oc.trk.optimize_scale(1, 1.0, oc.levels - 1)
    ts_o = oc.trk.trace()
    got = {}
    for depth in ("1", "4"):
        monkeypatch.setenv("DSLAM_LM_SPEC", depth)
        s = api.Session(0)
        gc = GpuCase(s, oc, template="device")
        l0 = s.launch_count()
        ok, pose, aff, last = gc.trk.trackNewestCoarse(gc.f_new, init, (0, 0), oc.levels - 1)
        n_track = s.launch_count() - l0
        tg = gc.trk.trace()
        l0 = s.launch_count()
        rmse, sc = gc.trk.optimizeScale(gc.f_right, 1.0, oc.levels - 1)
        n_scale = s.launch_count() - l0
        tsg = gc.trk.trace()
        assert ok == ok_o and _compare_traces(tg, to, 8) < 1e-5 and rel_err(pose, pose_o) < 1e-8
        assert tsg.shape == ts_o.shape and np.array_equal(tsg[:, :4], ts_o[:, :4]) and abs(sc - s_o) <= 1e-6 * abs(s_o)
        got[depth] = (n_track, n_scale, tg, tsg)
        gc.close()
        s.close()
    assert np.array_equal(got["1"][2][:, :4], got["4"][2][:, :4]) and np.allclose(got["1"][2], got["4"][2], rtol=1e-12, atol=0, equal_nan=True)
    assert got["4"][0] < got["1"][0] and got["4"][1] < got["1"][1], (got["1"][:2], got["4"][:2])
    print("launches track / scale: no speculation %d / %d, depth 4: %d / %d" % (got["1"][0], got["1"][1], got["4"][0], got["4"][1]))
