#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/dslam_oracle.cpp, -O2 -ffp-contract=off).

The reference ships no tests or fixtures for this path and cannot be built here, so these vectors pin
the ORACLE (and through it the CUDA path) against regressions and across machines; the pieces of the reference that do
compile in place (Accumulator9, ScaleAccumulator, search_place.h -> oracle/_ref) are checked separately in
tests/test_oracle_ref.py.  Inputs are stored in the fixture (images quantised to 1/64 grey level so they are exactly
representable) because numpy's transcendental functions are not bit-stable across platforms.

    python tests/golden/make_golden.py        # rewrites tracking_tiny.npz, scan_context_small.npz
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from direct_stereo_slam_b200 import synthetic as syn  # noqa: E402

IDENT7 = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


def quant(img):
    return np.round(np.asarray(img, np.float64) * 64).astype(np.uint16)


def dequant(q):
    return (q.astype(np.float32) / np.float32(64)).astype(np.float32)


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def tracking():
    o = orc.Oracle()
    c = syn.make_tracking_case("tiny", 3, scale_error=1.7)
    cfg = c["cfg"]
    w, h = cfg["w"], cfg["h"]
    levels = orc.pyr_levels_used(w, h)
    K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
    T = syn.t_stereo(cfg)
    q = {k: quant(c[k]) for k in ("img_ref", "img_new", "img_right")}
    img = {k: dequant(v) for k, v in q.items()}
    out = dict(w=w, h=h, levels=levels, K=K, T_stereo=T, pu=c["pu"], pv=c["pv"], pid=c["pid"], pw=c["pw"], pose7_true=c["pose7_true"], **{k + "_q64": v for k, v in q.items()})
    B = (255.0 * (np.arange(256) / 255.0) ** 0.8).astype(np.float32)
    out["B256"] = B
    off = orc.level_offsets(w, h, levels)
    pyr = {}
    for k in img:
        dIp, ag = o.make_images(img[k], levels, B256=B if k == "img_ref" else None)
        pyr[k] = dIp
        # checksums over the rows the reference defines (1 .. h-2) + a few raw samples
        crcs = []
        for l in range(levels):
            wl, hl = w >> l, h >> l
            d = dIp[off[l]:off[l + 1]].reshape(hl, wl, 3)
            a = ag[off[l]:off[l + 1]].reshape(hl, wl)
            crcs.append([crc(d[..., 0]), crc(d[1:-1, :, 1:]), crc(a[1:-1])])
        out["pyr_crc_" + k] = np.array(crcs, np.uint32)
        out["pyr_sample_" + k] = dIp[::997].copy()
    trk = o.tracker(w, h, levels, K, K, T)
    trk.make_coarse_depth(c["pu"], c["pv"], c["pid"], c["pw"], pyr["img_ref"])
    trk.set_ref_aff(1.0, 0.0, 0.0)
    trk.set_new_frame(pyr["img_new"], 1.0)
    trk.set_right_frame(pyr["img_right"])
    for l in range(levels):
        u, v, idp, col = trk.get_ref_level(l)
        out["pc_n_%d" % l] = np.int32(len(u))
        out["pc_crc_%d" % l] = np.array([crc(u), crc(v), crc(idp), crc(col)], np.uint32)
    # single evaluations
    rng = np.random.default_rng(5)
    poses = [IDENT7, c["pose7_true"]] + [o.se3_mul(o.se3_exp(np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.002, 3)])), c["pose7_true"]) for _ in range(3)]
    affs = [(0.0, 0.0), (0.03, 4.0)] + [tuple(rng.normal(0, [0.05, 5.0])) for _ in range(3)]
    ev = []
    for lvl in range(levels):
        for pose, aff in zip(poses, affs):
            for mode in (0, 1):
                trk.set_res_acc_mode(mode)
                res, n = trk.calc_res_pose(lvl, pose, aff, 20.0)
                H, b, acc = trk.calc_gs_pose(lvl, mode, aff)
                ev.append(np.concatenate([[lvl, mode, n], pose, aff, res, acc, H.reshape(-1), b]))
    out["pose_evals"] = np.array(ev)
    sev = []
    for lvl in range(levels):
        for s in (0.5, 1.0, 1.7, 4.0):
            for mode in (0, 1):
                trk.set_res_acc_mode(mode)
                res, n = trk.calc_res_scale(lvl, s, 20.0)
                H, b, acc = trk.calc_gs_scale(lvl, mode, s)
                sev.append(np.concatenate([[lvl, mode, n, s], res, acc, [H, b]]))
    out["scale_evals"] = np.array(sev)
    for mode in (0, 1):
        ok, pose, aff, last, flow = trk.track_newest_coarse(mode, IDENT7, (0, 0), levels - 1)
        out["track_trace_%d" % mode] = trk.trace()
        out["track_result_%d" % mode] = np.concatenate([[ok], pose, aff, last, flow])
        rmse, s = trk.optimize_scale(mode, 1.0, levels - 1)
        out["scale_trace_%d" % mode] = trk.trace()
        out["scale_result_%d" % mode] = np.array([rmse, s])
    np.savez_compressed(os.path.join(HERE, "tracking_tiny.npz"), **out)
    print("tracking_tiny.npz: pose result", out["track_result_1"][1:8], "scale", out["scale_result_1"])


def scan_context():
    o = orc.Oracle()
    rng = np.random.default_rng(7)
    n = 60
    ptr, idxs, vals, keys = [0], [], [], []
    clouds = []
    for r in range(n):
        pts = np.clip(np.round(rng.normal(0, 12, (1500, 3)) * 256), -32767, 32767) / 256  # exactly representable coordinates (int16 / 256)
        clouds.append(pts)
        rk, si, sv, tfm = o.sc_generate(pts)
        keys.append(rk)
        idxs.append(si)
        vals.append(sv)
        ptr.append(ptr[-1] + len(si))
    keys = np.stack(keys)
    idxs, vals = np.concatenate(idxs), np.concatenate(vals)
    # queries: revisits (jittered clouds) and strangers
    q_res = []
    q_sig = []
    for qn in range(12):
        if qn % 2 == 0:
            src = int(rng.integers(0, n))
            pts = clouds[src] + np.round(rng.normal(0, 0.2, clouds[src].shape) * 256) / 256
        else:
            pts = np.round(rng.normal(0, 12, (1500, 3)) * 256) / 256
        rk, si, sv, _ = o.sc_generate(pts)
        cand, dist = o.search_ringkey(rk, keys, k=3, thres=0.1)
        if len(cand) == 0:
            cand = np.array([0, 1, 2], np.int32)
        ri, rd = o.search_sc(si, sv, ptr, idxs, vals, cand)
        dense = np.zeros(1200, np.float64)
        dense[si] = sv
        q_sig.append(dense)
        q_res.append(np.concatenate([[len(cand)], np.pad(cand, (0, 3 - len(cand)), constant_values=-1), [ri, rd], rk]))
    np.savez_compressed(os.path.join(HERE, "scan_context_small.npz"), clouds_q256=np.round(np.stack(clouds) * 256).astype(np.int16), keys=keys, sig_ptr=np.array(ptr, np.int32), sig_idx=idxs,
                        sig_val=vals, q_sig=np.stack(q_sig), q_res=np.array(q_res))
    print("scan_context_small.npz:", n, "descriptors,", len(q_res), "queries")


if __name__ == "__main__":
    tracking()
    scan_context()
