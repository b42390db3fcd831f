"""One batched pose / scale evaluation at a given batch size (for ncu captures): python tools/one_eval.py kitti 128"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from direct_stereo_slam_b200 import api, synthetic as syn  # noqa: E402

cfg_name, nb = sys.argv[1], int(sys.argv[2])
c = syn.make_tracking_case(cfg_name, 42)
cfg = c["cfg"]
s = api.Session(0)
w, h = cfg["w"], cfg["h"]
levels = api.pyr_levels_used(w, h)
K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
ref, new, right = (api.FrameHessian(s, w, h, levels) for _ in range(3))
ref.makeImages(c["img_ref"], host=False)
new.makeImages(c["img_new"], host=False)
right.makeImages(c["img_right"], host=False)
trk = api.TrackerAndScaler(s, w, h, syn.t_stereo(cfg).reshape(-1), K, K0=K, levels=levels)
trk.setCoarseTrackingRef(ref, c["pu"], c["pv"], c["pid"], c["pw"])
rng = np.random.default_rng(0)
poses = np.tile(c["pose7_true"], (nb, 1))
poses[:, 4:] += rng.normal(0, 0.01, (nb, 3))
affs = rng.normal(0, [0.01, 1.0], (nb, 2))
for _ in range(4):
    trk.calcResAndGSPose(new, 0, poses, affs)
    trk.calcResAndGSScale(right, 0, rng.uniform(0.5, 2.0, nb).astype(np.float32))
s.sync()
