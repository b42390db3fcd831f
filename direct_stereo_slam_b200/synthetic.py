"""Deterministic synthetic stereo scenes standing in for KITTI / Malaga (no datasets on disk).

A scene is analytic geometry in the reference (left keyframe) camera frame — a ground plane, a far wall and a
set of fronto-parallel boxes — carrying a band-limited 3-D texture (sum of random sinusoids of the world
coordinates).  Any pinhole view (left keyframe, right camera through T_stereo, a new left frame after a
rigid motion) is rendered by exact ray casting, so all views are photometrically consistent and true depth
is known per pixel.  Calibrations come from the reference's cams/ directory:
  KITTI  : cams/kitti/0_2/camera0.txt:1-4  (f=718.856, c=(607.1928,185.2157), crop 1232x368), T_stereo tx=-0.5372
  Malaga : cams/malaga/camera0.txt:1-4     (f=795.11588, c=(517.12973,395.59665), 1024x768), T_stereo tx=-0.119471
"""
import numpy as np

CONFIGS = {
    "kitti": dict(w=1232, h=368, fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, baseline=0.5372, npts=2000),
    "malaga": dict(w=1024, h=768, fx=795.11588, fy=795.11588, cx=517.12973, cy=395.59665, baseline=0.119471, npts=2000),
    "synth1920": dict(w=1920, h=1200, fx=1400.0, fy=1400.0, cx=959.5, cy=599.5, baseline=0.3, npts=8000),
    # small sizes for fast CPU tests (same rule for pyramid levels: halve while even and area > 5000)
    "tiny": dict(w=320, h=192, fx=200.0, fy=200.0, cx=159.5, cy=95.5, baseline=0.3, npts=600),
}


def t_stereo(cfg):
    """Row-major 4x4 T_f1_f0 like cams/*/T_stereo.yaml (tz = 1e-9 as shipped)."""
    T = np.eye(4)
    T[0, 3] = -cfg["baseline"]
    T[2, 3] = 1e-9
    return T


def so3_exp(om):
    th = np.linalg.norm(om)
    K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


def se3_exp_mat(xi):
    """xi = (upsilon, omega) -> (R, t) with the usual SE3 exponential."""
    ups, om = np.asarray(xi[:3], float), np.asarray(xi[3:], float)
    th = np.linalg.norm(om)
    K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    R = so3_exp(om)
    if th < 1e-12:
        V = np.eye(3) + 0.5 * K
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * (K @ K)
    return R, V @ ups


def R_to_quat(R):
    """(x, y, z, w) — the layout of Sophus::SE3d::data()."""
    tr = np.trace(R)
    if tr > 0:
        t = np.sqrt(tr + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        return np.array([(R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t, w])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[i] = 0.5 * t
    t = 0.5 / t
    q[3] = (R[k, j] - R[j, k]) * t
    q[j] = (R[j, i] + R[i, j]) * t
    q[k] = (R[k, i] + R[i, k]) * t
    return q


def pose7(R, t):
    return np.concatenate([R_to_quat(R), np.asarray(t, float)])


def quat_mul(a, b):
    """Hamilton product of two (x, y, z, w) quaternions."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


ROT_SIGNS = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (-1, 1, 0), (0, -1, 1), (-1, 0, 1),
             (1, -1, 0), (0, 1, -1), (1, 0, -1), (-1, -1, 0), (0, -1, -1), (-1, 0, -1), (-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1),
             (1, -1, 1), (1, 1, -1), (1, 1, 1)]


def frontend_pose_tries(const_motion7, double_motion7, half_motion7, zero_motion7):
    """The 83 initialisations FrontEnd::trackNewCoarse tries in order (src/FrontEnd.cpp:147-180): constant / double / half /
    zero motion, zero motion from the keyframe, then constant motion composed with 26 sign patterns x 3 small rotations
    (rot_delta = 0.02, 0.03, 0.04; the float loop `< 0.05` ends there).  Poses are (qx, qy, qz, qw, tx, ty, tz)."""
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    tries = [np.asarray(const_motion7, np.float64), np.asarray(double_motion7, np.float64), np.asarray(half_motion7, np.float64),
             np.asarray(zero_motion7, np.float64), ident]
    base = tries[0]
    for d in (0.02, 0.03, 0.04):
        for sgn in ROT_SIGNS:
            q = np.array([sgn[0] * d, sgn[1] * d, sgn[2] * d, 1.0])
            q /= np.linalg.norm(q)  # the SE3(Quaterniond, t) constructor normalises
            tries.append(np.concatenate([quat_mul(base[:4], q), base[4:]]))  # base * SE3(q, 0): rotation composes, translation stays
    return np.stack(tries)


class Scene:
    def __init__(self, cfg, seed, n_waves=32, n_boxes=12):
        self.cfg = cfg
        rng = np.random.default_rng(seed)
        self.seed = seed
        # texture: sum of sinusoids of world coordinates, wavelengths 0.25 .. 5 m
        lam = np.exp(rng.uniform(np.log(0.25), np.log(5.0), n_waves))
        dirs = rng.normal(size=(n_waves, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        self.kvec = (2 * np.pi / lam)[:, None] * dirs
        self.phase = rng.uniform(0, 2 * np.pi, n_waves)
        self.amp = rng.uniform(0.5, 1.0, n_waves) * 70.0 / np.sqrt(n_waves)
        # geometry (reference camera frame, x right, y down, z forward)
        self.ground_y = 1.65
        self.wall_z = 60.0
        boxes = []
        for _ in range(n_boxes):
            z = rng.uniform(4.0, 40.0)
            xc = rng.uniform(-0.6, 0.6) * z
            wx = rng.uniform(0.8, 3.0)
            hy = rng.uniform(0.8, 3.0)
            boxes.append((z, xc - wx / 2, xc + wx / 2, self.ground_y - hy, self.ground_y, rng.uniform(-30, 30)))
        self.boxes = boxes

    def _texture(self, P, offset):
        acc = np.full(P.shape[:-1], 128.0) + offset
        for k, ph, a in zip(self.kvec, self.phase, self.amp):
            acc += a * np.sin(P @ k + ph)
        return acc

    def render(self, R=None, t=None, K=None, noise_seed=None, noise_sigma=2.0, aff=(0.0, 0.0)):
        """Render the view of camera X_cam = R X_ref + t.  Returns (image float32 [h,w] in 0..255, depth [h,w])."""
        cfg = self.cfg
        w, h = cfg["w"], cfg["h"]
        R = np.eye(3) if R is None else np.asarray(R, float)
        t = np.zeros(3) if t is None else np.asarray(t, float)
        fx, fy, cx, cy = (cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]) if K is None else K
        uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        dcam = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1)
        d = dcam @ R  # R^T d_cam, row-vector form
        c = -R.T @ t
        best_s = np.full((h, w), np.inf)
        off = np.zeros((h, w))
        # far wall
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (self.wall_z - c[2]) / d[..., 2]
            ok = s > 0
            best_s = np.where(ok, s, best_s)
            # ground
            s = (self.ground_y - c[1]) / d[..., 1]
            ok = (s > 0) & (s < best_s)
            best_s = np.where(ok, s, best_s)
            off = np.where(ok, 10.0, off)
            for (z, x0, x1, y0, y1, o) in self.boxes:
                s = (z - c[2]) / d[..., 2]
                X = c[0] + s * d[..., 0]
                Y = c[1] + s * d[..., 1]
                ok = (s > 0) & (s < best_s) & (X >= x0) & (X <= x1) & (Y >= y0) & (Y <= y1)
                best_s = np.where(ok, s, best_s)
                off = np.where(ok, o, off)
        best_s = np.where(np.isfinite(best_s), best_s, self.wall_z)
        P = c[None, None, :] + best_s[..., None] * d
        img = self._texture(P, off)
        img = np.exp(aff[0]) * img + aff[1]
        if noise_sigma > 0:
            rng = np.random.default_rng(self.seed * 7919 + 13 if noise_seed is None else noise_seed)
            img = img + rng.normal(0, noise_sigma, img.shape)
        img = np.clip(img, 0, 255).astype(np.float32)
        depth = best_s * dcam[..., 2]  # z in the rendered camera = s * (R d)_z = s (dcam_z = 1)
        return img, depth.astype(np.float64)


def select_points(img, depth, npts, seed, grad_thresh=8.0, scale_error=1.0):
    """Template candidates like DSO's active points: pixels with gradient magnitude > thresh inside a 3-px margin.
    Returns integer pixel (u, v), idepth (true inverse depth * scale_error) and weight (1)."""
    h, w = img.shape
    gx = np.zeros_like(img)
    gy = np.zeros_like(img)
    gx[:, 1:-1] = 0.5 * (img[:, 2:] - img[:, :-2])
    gy[1:-1, :] = 0.5 * (img[2:, :] - img[:-2, :])
    g = np.sqrt(gx * gx + gy * gy)
    mask = g > grad_thresh
    mask[:3, :] = False
    mask[-4:, :] = False
    mask[:, :3] = False
    mask[:, -4:] = False
    vs, us = np.nonzero(mask)
    rng = np.random.default_rng(seed)
    if len(us) > npts:
        sel = rng.choice(len(us), npts, replace=False)
        sel.sort()
        us, vs = us[sel], vs[sel]
    idepth = (1.0 / depth[vs, us]) * scale_error
    return us.astype(np.int32), vs.astype(np.int32), idepth.astype(np.float32), np.ones(len(us), np.float32)


def make_tracking_case(cfg_name, seed, motion_scale=1.0, scale_error=1.0, n_waves=32, with_right=True):
    """One synthetic tracking problem: keyframe image + template candidates, a new
    left frame after the motion xi* (scaled), the right image of the keyframe, ground-truth pose and affine."""
    cfg = CONFIGS[cfg_name]
    sc = Scene(cfg, seed, n_waves=n_waves)
    rng = np.random.default_rng(seed + 100003)
    img_ref, depth_ref = sc.render(noise_seed=seed * 3 + 1)
    m = motion_scale * rng.uniform(0.5, 1.5)
    xi = np.array([0.02, -0.01, 0.35, 0.004, -0.010, 0.002]) * m
    R, t = se3_exp_mat(xi)
    aff_true = (0.03, 4.0)
    img_new, _ = sc.render(R, t, noise_seed=seed * 3 + 2, aff=aff_true)
    out = dict(cfg=cfg, scene=sc, img_ref=img_ref, depth_ref=depth_ref, img_new=img_new, xi_true=xi, R_true=R, t_true=t,
               pose7_true=pose7(R, t), aff_true=aff_true)
    if with_right:
        Ts = t_stereo(cfg)
        img_right, _ = sc.render(Ts[:3, :3], Ts[:3, 3], noise_seed=seed * 3 + 3)
        out["img_right"] = img_right
    pu, pv, pid, pw = select_points(img_ref, depth_ref, cfg["npts"], seed + 5, scale_error=scale_error)
    out.update(pu=pu, pv=pv, pid=pid, pw=pw, scale_error=scale_error)
    return out


def make_sc_database(n, seed, num_s=60, num_r=20, empty_frac=0.35):
    """Dense synthetic Scan-Context DB: heights N(0,1), `empty_frac` of the cells empty (0),
    each sector column L2-normalised; ring key = occupied fraction per ring.  Layout [n, num_s*num_r], cell index
    = sector*num_r + ring like ScanContext.cpp:119."""
    rng = np.random.default_rng(seed)
    sig = rng.normal(size=(n, num_s, num_r)).astype(np.float32)
    occ = rng.random((n, num_s, num_r)) >= empty_frac
    sig = np.where(occ, sig, 0).astype(np.float32)
    nrm = np.sqrt((sig.astype(np.float64) ** 2).sum(-1, keepdims=True))
    nrm[nrm == 0] = 1
    sig = (sig / nrm).astype(np.float32)
    ringkey = (occ.sum(1) / float(num_s)).astype(np.float32)
    return sig.reshape(n, num_s * num_r), ringkey


def make_sc_queries(db_sig, db_key, q, seed, noise=0.05, num_s=60, num_r=20):
    """Half the queries are DB rows + N(0, noise) on occupied cells (known answer), half are fresh random."""
    rng = np.random.default_rng(seed)
    n = db_sig.shape[0]
    qs = np.empty((q, db_sig.shape[1]), np.float32)
    qk = np.empty((q, num_r), np.float32)
    truth = np.full(q, -1, np.int64)
    fresh_sig, fresh_key = make_sc_database(q, seed + 1, num_s, num_r)
    for i in range(q):
        if i % 2 == 0:
            r = int(rng.integers(0, n))
            s = db_sig[r].astype(np.float64).reshape(num_s, num_r)
            occ = s != 0
            s = s + occ * rng.normal(0, noise, s.shape) * np.abs(s).max()
            nrm = np.sqrt((s**2).sum(-1, keepdims=True))
            nrm[nrm == 0] = 1
            qs[i] = (s / nrm).reshape(-1).astype(np.float32)
            qk[i] = db_key[r]
            truth[i] = r
        else:
            qs[i] = fresh_sig[i]
            qk[i] = fresh_key[i]
    return qs, qk, truth
