"""include/dslam_b200_adapter.hpp instantiated with the REFERENCE'S OWN types (dso::FrameHessian, SE3, AffLight, Vec5, Vec3,
CalibHessian — provided in this image by the stand-ins of oracle/shim) and run next to the reference's own TrackerAndScaler.cpp
through the same FrontEnd-style call sequence (src/FrontEnd.cpp:57-58, 204-206, 605, 680, 797-798, 992, 1032 and the hypothesis
loop :192-247).  The program is oracle/adapter_vs_reference.cpp, built in place by oracle/ref_build.py into
oracle/_ref/adapter_vs_reference (the binary travels to the GPU box; /root/reference is not needed at run time)."""
import os
import subprocess

import numpy as np
import pytest

from direct_stereo_slam_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "adapter_vs_reference")


def _write_scene(path, cfg_name, seed):
    c = syn.make_tracking_case(cfg_name, seed, scale_error=1.3)
    cfg = c["cfg"]
    w, h = cfg["w"], cfg["h"]
    levels = 1
    ww, hh = w, h
    while ww % 2 == 0 and hh % 2 == 0 and ww * hh > 5000 and levels < 6:
        ww, hh, levels = ww // 2, hh // 2, levels + 1
    R2, t2 = syn.se3_exp_mat(c["xi_true"] * 0.6)
    img_new2, _ = c["scene"].render(R2, t2, noise_seed=seed * 3 + 7, aff=c["aff_true"])
    init = np.stack([syn.pose7(*syn.se3_exp_mat(c["xi_true"] * 0.9)), syn.pose7(*syn.se3_exp_mat(c["xi_true"] * 0.6 * 0.9))])
    hdif = np.full(len(c["pu"]), 1e-3, np.float32)  # weight = sqrtf(1e-3 / (HdiF + 1e-12)) ~ 1
    with open(path, "wb") as f:
        np.array([w, h, levels, len(c["pu"])], np.int32).tofile(f)
        np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32).tofile(f)
        syn.t_stereo(cfg).astype(np.float64).tofile(f)
        init.astype(np.float64).tofile(f)
        for img in (c["img_ref"], c["img_new"], img_new2, c["img_right"]):
            np.ascontiguousarray(img, np.float32).tofile(f)
        np.ascontiguousarray(c["pu"], np.int32).tofile(f)
        np.ascontiguousarray(c["pv"], np.int32).tofile(f)
        np.ascontiguousarray(c["pid"], np.float32).tofile(f)
        hdif.tofile(f)


def test_adapter_vs_reference_is_built_against_the_reference_types():
    """CPU part: the program exists whenever oracle/_ref was built (ref_build.py compiles it: the adapter templates instantiate
    with the reference's types) and degrades to exit code 3 without a CUDA device."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/adapter_vs_reference not built (needs /root/reference at build time)")
    r = subprocess.run([EXE, "/nonexistent"], capture_output=True, text=True)
    assert r.returncode in (2, 3), r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_name,seed", [("tiny", 3), ("kitti", 1000)])
def test_adapter_equals_reference_tracker(tmp_path, cfg_name, seed):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/adapter_vs_reference did not travel")
    scene = tmp_path / "scene.bin"
    _write_scene(scene, cfg_name, seed)
    r = subprocess.run([EXE, str(scene)], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout
