"""Pyramid build timing (CUDA events on the session stream): python tools/pyramid_bench.py [nframes] [kitti|malaga|synth1920]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from direct_stereo_slam_b200 import api, synthetic as syn

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = syn.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "kitti"]
w, h = cfg["w"], cfg["h"]
levels = api.pyr_levels_used(w, h)
s = api.Session(0)
rng = np.random.default_rng(0)
frames = [api.FrameHessian(s, w, h, levels) for _ in range(nf)]
img = np.clip(rng.normal(128, 40, (h, w)), 0, 255).astype(np.float32)
for f in frames:
    f.upload(img)
tot = sum((w >> l) * (h >> l) for l in range(levels))
for stage, name in ((0, "texels only"), (3, "texels + host-layout staging")):
    alg = nf * (4 * w * h + 16 * tot + (16 * tot if stage else 0))
    for _ in range(3):
        api.build_frames(frames, stage_host=stage)
    ts = []
    for _ in range(10):
        s.mark(0)
        api.build_frames(frames, stage_host=stage)
        s.mark(1)
        ts.append(s.elapsed_ms())
    t = float(np.median(ts))
    print("%s %dx%d x%d frames, %s: %.3f ms  (%.1f us/frame)  %.0f GB/s algorithmic (%.1f MB/frame)" % (sys.argv[2] if len(sys.argv) > 2 else "kitti", w, h, nf, name, t, t * 1e3 / nf, alg / t / 1e6, alg / nf / 1e6))
