"""Per-launch timing of the fused pose / scale residual kernels vs batch size (CUDA events around every launch).
Usage (on the GPU box): python tools/kernel_sweep.py [kitti|synth1920] """
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from direct_stereo_slam_b200 import api, synthetic as syn  # noqa: E402


def main():
    cfg_name = sys.argv[1] if len(sys.argv) > 1 else "kitti"
    c = syn.make_tracking_case(cfg_name, 42)
    cfg = c["cfg"]
    s = api.Session(0)
    w, h = cfg["w"], cfg["h"]
    levels = api.pyr_levels_used(w, h)
    K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
    ref = api.FrameHessian(s, w, h, levels)
    new = api.FrameHessian(s, w, h, levels)
    right = api.FrameHessian(s, w, h, levels)
    ref.makeImages(c["img_ref"], host=False)
    new.makeImages(c["img_new"], host=False)
    right.makeImages(c["img_right"], host=False)
    trk = api.TrackerAndScaler(s, w, h, syn.t_stereo(cfg).reshape(-1), K, K0=K, levels=levels)
    pcn = trk.setCoarseTrackingRef(ref, c["pu"], c["pv"], c["pid"], c["pw"])
    print("config %s %dx%d levels %d pc_n %s" % (cfg_name, w, h, levels, pcn.tolist()))
    rng = np.random.default_rng(0)
    for lvl in (0, 2):
        for nb in (1, 8, 32, 128, 512, 1024):
            poses = np.tile(c["pose7_true"], (nb, 1))
            poses[:, 4:] += rng.normal(0, 0.01, (nb, 3))
            affs = rng.normal(0, [0.01, 1.0], (nb, 2))
            for _ in range(3):
                trk.calcResAndGSPose(new, lvl, poses, affs)
            s.profile(True)
            for _ in range(10):
                trk.calcResAndGSPose(new, lvl, poses, affs)
            p = s.profile_read()["pose"]
            s.profile(False)
            us = p["ms"] * 1e3 / p["launches"]
            pts = p["points"] / p["launches"]
            print("pose  lvl %d nb %4d: %6.1f launches/call  %8.2f us/launch  %9.0f pts/launch  %7.1f GB/s (64 B/pt)  %6.2f ns/kpt" % (
                lvl, nb, p["launches"] / 10, us, pts, 64 * pts / us / 1e3, us * 1e3 / pts * 1e3 / 1e3))
    for nb in (1, 8, 128, 1024):
        scales = rng.uniform(0.5, 2.0, nb).astype(np.float32)
        for _ in range(3):
            trk.calcResAndGSScale(right, 0, scales)
        s.profile(True)
        for _ in range(10):
            trk.calcResAndGSScale(right, 0, scales)
        p = s.profile_read()["scale"]
        s.profile(False)
        us = p["ms"] * 1e3 / p["launches"]
        pts = p["points"] / p["launches"]
        print("scale lvl 0 nb %4d: %8.2f us/launch  %9.0f pts/launch  %7.1f GB/s" % (nb, us, pts, 64 * pts / us / 1e3))


if __name__ == "__main__":
    main()
