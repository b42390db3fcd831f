"""The CUDA path against the COMMITTED golden vectors (tests/golden/*.npz) — no oracle call on the checked side."""
import os
import zlib

import numpy as np
import pytest

from direct_stereo_slam_b200 import api

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


@pytest.fixture(scope="module")
def setup(session):
    g = np.load(os.path.join(G, "tracking_tiny.npz"))
    w, h, levels = int(g["w"]), int(g["h"]), int(g["levels"])
    fr = {}
    for k in ("img_ref", "img_new", "img_right"):
        f = api.FrameHessian(session, w, h, levels)
        f.makeImages(g[k + "_q64"].astype(np.float32) / np.float32(64), B256=g["B256"] if k == "img_ref" else None)
        fr[k] = f
    trk = api.TrackerAndScaler(session, w, h, g["T_stereo"].reshape(-1), g["K"], K0=g["K"], levels=levels)
    pcn = trk.setCoarseTrackingRef(fr["img_ref"], g["pu"], g["pv"], g["pid"], g["pw"])
    yield g, w, h, levels, fr, trk, pcn
    trk.close()
    for f in fr.values():
        f.close()


def test_pyramid_golden(setup):
    g, w, h, levels, fr, trk, pcn = setup
    for k, f in fr.items():
        for l in range(levels):
            d, a = f.dIp(l), f.absSquaredGrad(l)
            assert [crc(d[..., 0]), crc(d[1:-1, :, 1:]), crc(a[1:-1])] == list(g["pyr_crc_" + k][l]), (k, l)
        assert np.array_equal(f.dIp_all[::997][:, 0], g["pyr_sample_" + k][:, 0])


def test_template_golden(setup):
    g, w, h, levels, fr, trk, pcn = setup
    for l in range(levels):
        u, v, idp, col = trk.ref_level(l)
        assert len(u) == int(g["pc_n_%d" % l]) == pcn[l]
        assert [crc(u), crc(v), crc(idp), crc(col)] == list(g["pc_crc_%d" % l])


def test_evaluations_golden(setup):
    g, w, h, levels, fr, trk, pcn = setup
    for row in g["pose_evals"]:
        lvl, mode, n = int(row[0]), int(row[1]), int(row[2])
        if mode != 1:
            continue
        out = trk.calcResAndGSPose(fr["img_new"], lvl, row[3:10], row[10:12], 20.0)
        res, acc, H, b = row[12:18], row[18:63], row[63:127], row[127:135]
        assert out["n"][0] == n and out["res6"][0, 1] == res[1] and out["res6"][0, 5] == res[5]
        assert np.allclose(out["res6"][0], res, rtol=1e-11, atol=0)
        assert np.allclose(out["acc48"][0, :45], acc, rtol=1e-9, atol=1e-9 * np.abs(acc).max())
        assert np.linalg.norm(out["H"][0].reshape(-1) - H) < 1e-11 * np.linalg.norm(H)
        assert np.linalg.norm(out["b"][0] - b) < 1e-10 * np.linalg.norm(b)
    for row in g["scale_evals"]:
        lvl, mode, n, s = int(row[0]), int(row[1]), int(row[2]), float(row[3])
        if mode != 1:
            continue
        out = trk.calcResAndGSScale(fr["img_right"], lvl, s, 20.0)
        assert out["n"][0] == n
        assert np.allclose(out["res6"][0], row[4:10], rtol=1e-11, atol=0)
        assert np.allclose(out["acc8"][0, :3], row[10:13], rtol=1e-10, atol=1e-12)


def test_lm_traces_golden(setup):
    g, w, h, levels, fr, trk, pcn = setup
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    ok, pose, aff, last = trk.trackNewestCoarse(fr["img_new"], ident, (0, 0), levels - 1)
    tg, to = trk.trace(), g["track_trace_1"]
    assert tg.shape == to.shape and np.array_equal(tg[:, :4], to[:, :4])
    for rg, ro in zip(tg, to):
        if ro[1] >= 0:
            assert np.linalg.norm(rg[7:] - ro[7:]) <= 1e-5 * np.linalg.norm(ro[7:])  # north-star tolerance on the pose delta
    want = g["track_result_1"]
    assert bool(want[0]) == ok and np.allclose(pose, want[1:8], rtol=1e-8, atol=1e-11) and np.allclose(aff, want[8:10], rtol=1e-7)
    rmse, s = trk.optimizeScale(fr["img_right"], 1.0, levels - 1)
    ts, tso = trk.trace(), g["scale_trace_1"]
    assert ts.shape == tso.shape and np.array_equal(ts[:, :4], tso[:, :4])
    assert np.allclose([rmse, s], g["scale_result_1"], rtol=1e-6)


def test_scan_context_golden(session):
    g = np.load(os.path.join(G, "scan_context_small.npz"))
    ptr, sidx, sval = g["sig_ptr"], g["sig_idx"], g["sig_val"]
    n = len(ptr) - 1
    db = api.ScanContextDB(session, 64)
    for r in range(n):
        db.add_sparse(g["keys"][r], sidx[ptr[r]:ptr[r + 1]], sval[ptr[r]:ptr[r + 1]])
    for qd, row in zip(g["q_sig"], g["q_res"]):
        ncand = int(row[0])
        cand = row[1:4].astype(np.int32)
        rk = row[6:].astype(np.float32)
        c2, d2 = db.search_ringkey(rk, k=3, thres=0.1)
        got = c2[0][c2[0] >= 0]
        if len(got):
            assert np.array_equal(got, cand[:ncand])
        i, d = db.search_sc(qd.astype(np.float32), cand)
        # the database stores fp32 (the fixture's expectation used the double values): same winner, distance to ~1e-7
        assert i[0] == int(row[4]) and abs(d[0] - row[5]) < 2e-7
    db.close()
