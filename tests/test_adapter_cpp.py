"""include/dslam_b200_adapter.hpp (the C++ glue for the reference's call sites) compiles against mock DSO types and links
against the in-tree library; with a GPU the mock program tracks a shifted textured plane through the adapter."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_mock")


def _build():
    cmd = ["g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adapter_mock.cpp"),
           "-L", os.path.join(ROOT, "direct_stereo_slam_b200"), "-ldslam_b200", "-Wl,-rpath," + os.path.join(ROOT, "direct_stereo_slam_b200"), "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_adapter_compiles_and_links():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapter_mock:" in r.stdout


@pytest.mark.gpu
def test_adapter_tracks_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok=1" in r.stdout, r.stdout
