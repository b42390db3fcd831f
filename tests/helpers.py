"""Shared builders for the parity tests: one synthetic tracking problem evaluated by the CPU oracle."""
import numpy as np

import oracle as orc
from direct_stereo_slam_b200 import synthetic as syn

IDENT7 = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


def cam_K(cfg):
    return np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)


class OracleCase:
    """Keyframe + new left frame + right frame of one synthetic scene, with the oracle's pyramids and template."""

    def __init__(self, o, cfg_name="tiny", seed=3, motion_scale=1.0, scale_error=1.0, levels=None):
        self.o = o
        self.case = c = syn.make_tracking_case(cfg_name, seed, motion_scale=motion_scale, scale_error=scale_error)
        cfg = self.cfg = c["cfg"]
        self.w, self.h = cfg["w"], cfg["h"]
        self.levels = orc.pyr_levels_used(self.w, self.h) if levels is None else levels
        self.K = cam_K(cfg)
        self.T_stereo = syn.t_stereo(cfg)
        self.dIp_ref, self.abs_ref = o.make_images(c["img_ref"], self.levels)
        self.dIp_new, self.abs_new = o.make_images(c["img_new"], self.levels)
        self.dIp_right, _ = o.make_images(c["img_right"], self.levels)
        self.trk = o.tracker(self.w, self.h, self.levels, self.K, self.K, self.T_stereo)
        self.trk.make_coarse_depth(c["pu"], c["pv"], c["pid"], c["pw"], self.dIp_ref)
        self.trk.set_ref_aff(1.0, 0.0, 0.0)
        self.trk.set_new_frame(self.dIp_new, 1.0)
        self.trk.set_right_frame(self.dIp_right)
        self.ref_levels = [self.trk.get_ref_level(l) for l in range(self.levels)]


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class GpuCase:
    """The device twin of an OracleCase: frames built by the CUDA pyramid, tracker fed with the oracle's template."""

    def __init__(self, session, oc, template="upload"):
        from direct_stereo_slam_b200 import api

        self.api = api
        self.oc = oc
        c = oc.case
        self.f_ref = api.FrameHessian(session, oc.w, oc.h, oc.levels)
        self.f_new = api.FrameHessian(session, oc.w, oc.h, oc.levels)
        self.f_right = api.FrameHessian(session, oc.w, oc.h, oc.levels)
        self.f_ref.makeImages(c["img_ref"], host=False)
        self.f_new.makeImages(c["img_new"], host=False)
        self.f_right.makeImages(c["img_right"], host=False)
        self.trk = api.TrackerAndScaler(session, oc.w, oc.h, oc.T_stereo.reshape(-1), oc.K, K0=oc.K, levels=oc.levels)
        if template == "upload":
            self.trk.setCoarseTrackingRefArrays(oc.ref_levels, ref_frame=self.f_ref)
        else:
            self.pc_n = self.trk.setCoarseTrackingRef(self.f_ref, c["pu"], c["pv"], c["pid"], c["pw"])

    def close(self):
        for x in (self.trk, self.f_ref, self.f_new, self.f_right):
            x.close()


def perturbed_pose(o, pose7, rng, trans=0.01, rot=0.002):
    xi = np.concatenate([rng.normal(0, trans, 3), rng.normal(0, rot, 3)])
    return o.se3_mul(o.se3_exp(xi), pose7)


def rotation_hypotheses(o, base7, deltas=(0.02, 0.03, 0.04)):
    """The rotation retries of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:164-180): base * SE3(Quaterniond(1, sx*d, sy*d, sz*d), 0)
    for 26 sign patterns and a few rotation steps (the quaternion is normalised by the SE3 constructor)."""
    signs = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (-1, 1, 0), (0, -1, 1), (-1, 0, 1),
             (1, -1, 0), (0, 1, -1), (1, 0, -1), (-1, -1, 0), (0, -1, -1), (-1, 0, -1), (-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1),
             (1, -1, 1), (1, 1, -1), (1, 1, 1)]
    out = []
    for d in deltas:
        for s in signs:
            q = np.array([s[0] * d, s[1] * d, s[2] * d, 1.0])
            q /= np.linalg.norm(q)
            out.append(o.se3_mul(base7, np.concatenate([q, [0, 0, 0]])))
    return np.stack(out)


def loop_closure_points(oc, n=1500, seed=0):
    """Reference points of a loop-closure alignment (LoopHandler::publishKeyframes, src/loop_closure/LoopHandler.cpp:166-178):
    3-D points of the keyframe in its camera frame + their colour on every pyramid level."""
    c = oc.case
    cfg = oc.cfg
    rng = np.random.default_rng(seed)
    sel = rng.choice(len(c["pu"]), min(n, len(c["pu"])), replace=False)
    u, v = c["pu"][sel].astype(np.float64), c["pv"][sel].astype(np.float64)
    z = 1.0 / c["pid"][sel].astype(np.float64) * c["scale_error"]
    pts = np.stack([(u - cfg["cx"]) / cfg["fx"] * z, (v - cfg["cy"]) / cfg["fy"] * z, z], 1)
    off = orc.level_offsets(oc.w, oc.h, oc.levels)
    colors = np.zeros((oc.levels, len(sel)), np.float32)
    for l in range(oc.levels):
        wl, hl = oc.w >> l, oc.h >> l
        I = oc.dIp_ref[off[l]:off[l + 1], 0].reshape(hl, wl)
        ul = np.clip(((u + 0.5) / (1 << l) - 0.5).round().astype(int), 0, wl - 1)
        vl = np.clip(((v + 0.5) / (1 << l) - 0.5).round().astype(int), 0, hl - 1)
        colors[l] = I[vl, ul]
    return pts, colors


def mat4_from_pose7(o, p7):
    T = np.eye(4)
    T[:3, :3] = o.se3_R(p7)
    T[:3, 3] = p7[4:]
    return T
