"""Where a bench step goes: pyramid-only, LM-only and full steps of bench.GpuStreams (device-resident inputs).
python tools/phase_split.py [streams] [steps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from direct_stereo_slam_b200 import api

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
session = api.Session(0)
cases = bench.make_cases(4)
st = bench.GpuStreams(api, session, cases, S)
st.upload_inputs()


def timed(fn, k0=3):
    for k in range(k0):
        fn(k)
    session.sync()
    t0 = time.perf_counter()
    for k in range(k0, k0 + K):
        fn(k)
    session.sync()
    return (time.perf_counter() - t0) * 1e3 / K


def pyr_only(k):
    v = k & 1
    left = [st.f_new[i][v] for i in range(st.n)]
    kf = [i for i in range(st.n) if st.is_kf(i, k)]
    api.build_frames(left)
    if kf:
        api.build_frames([st.f_right[i][v] for i in kf])


def lm_only(k):
    v = k & 1
    left = [st.f_new[i][v] for i in range(st.n)]
    kf = [i for i in range(st.n) if st.is_kf(i, k)]
    poses = np.stack([st.case_of[i]["pose_init"][v] for i in range(st.n)])
    api.lm_batch(st.trk, left, poses, np.zeros((st.n, 2)), st.levels - 1, [st.trk[i] for i in kf], [st.f_right[i][v] for i in kf], np.ones(len(kf), np.float32))


for k in range(2):  # all pyramids exist before the LM-only pass
    pyr_only(k)
full = timed(lambda k: st.step(k, False))
pyr = timed(pyr_only)
session.host_times()
lm = timed(lm_only)
ht = session.host_times()
print("S=%d groups=%s: full step %.3f ms (%.0f frames/s) | pyramids only %.3f ms | LM only %.3f ms | host per step (sum over threads): %s" % (
    S, os.environ.get("DSLAM_LM_GROUPS", "auto"), full, S / full * 1e3, pyr, lm, {k: round(v / (K + 3), 3) for k, v in ht.items()}))
