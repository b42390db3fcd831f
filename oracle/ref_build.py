#!/usr/bin/env python
"""Compile the pieces of the REFERENCE that build from their own few source files, in place, into oracle/_ref/.

What builds here (plain g++; the reference's cmake / catkin build is not run and cannot be — no Eigen, Boost, OpenCV, FLANN,
PCL, ROS in this image), every piece against stand-ins for the absent third-party headers (oracle/shim, oracle/shim_sc):
  * libdslam_ref.so            deps:dso/src/OptimizationBackend/MatrixAccumulators.h (Accumulator9), src/scale_optimization/
                               ScaleAccumulator.h, src/loop_closure/loop_detection/search_place.h            (ref_driver.cpp)
  * libdslam_ref_tracker.so    src/scale_optimization/TrackerAndScaler.cpp :1-336 and :451-1172 (constructor, makeK,
    libdslam_ref_tracker_opt   makeCoarseDepthL0, trackNewestCoarse, calcResPose, calcGSSSEPose, optimizeScale, calcResScale,
                               calcGSSSEScale) + FrameHessian::makeImages (deps:dso HessianBlocks.cpp:128-191); the _opt build
                               (-O3 -march=x86-64-v3) is the timed CPU baseline of bench.py               (ref_driver_tracker.cpp)
  * libdslam_ref_pe.so         src/loop_closure/pose_estimation/PoseEstimator.cpp, whole file                (ref_driver_pe.cpp)
  * libdslam_ref_sc.so         src/loop_closure/loop_detection/ScanContext.cpp, whole file                   (ref_driver_sc.cpp)
Files that live inside /root/reference/dependencies.zip are extracted to a temporary directory outside the repository for the
duration of the compile.  Outputs go to oracle/_ref/ only (git-ignored; they travel to the GPU box with the snapshot).  No
reference source is copied into the repository.
"""
import os
import subprocess
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DSLAM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(REF):
        print("reference tree %s not present: nothing to build (a prebuilt oracle/_ref is used if it travelled)" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="dslam_ref_") as tmp:
        with zipfile.ZipFile(os.path.join(REF, "dependencies.zip")) as z:
            z.extract("dso/src/OptimizationBackend/MatrixAccumulators.h", tmp)
        cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-msse2", "-ffp-contract=off", "-w",
               "-I", os.path.join(HERE, "shim"),          # Eigen/Core, util/NumType.h stand-ins (searched first)
               "-I", os.path.join(tmp, "dso", "src"),      # OptimizationBackend/MatrixAccumulators.h
               "-I", os.path.join(REF, "src"),             # scale_optimization/..., loop_closure/...
               os.path.join(HERE, "ref_driver.cpp"), "-o", os.path.join(OUT, "libdslam_ref.so")]
        subprocess.run(cmd, check=True)
        print("built", os.path.join(OUT, "libdslam_ref.so"))
        # ---- the tracker itself: the reference's TrackerAndScaler.cpp (hot-path line ranges) against the shims ----------
        with zipfile.ZipFile(os.path.join(REF, "dependencies.zip")) as z:
            z.extract("dso/src/util/globalFuncs.h", tmp)  # getInterpolatedElement33
            hb = z.read("dso/src/FullSystem/HessianBlocks.cpp").decode().split("\n")
        with open(os.path.join(tmp, "makeimages_extract.inc"), "w") as f:
            f.write("\n".join(hb[127:191]))  # FrameHessian::makeImages, :128-191
        src = open(os.path.join(REF, "src", "scale_optimization", "TrackerAndScaler.cpp")).read().split("\n")
        keep = src[0:336] + src[450:1172] + ["", "}  // namespace dso", ""]  # :1-336 and :451-1172 (1-based, inclusive)
        with open(os.path.join(tmp, "tracker_extract.inc"), "w") as f:
            f.write("\n".join(keep))
        cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-msse2", "-ffp-contract=off", "-w",
               "-I", os.path.join(HERE, "shim"), "-I", tmp, "-I", os.path.join(tmp, "dso", "src"),
               "-I", os.path.join(REF, "src"), "-I", os.path.join(REF, "src", "scale_optimization"),
               os.path.join(HERE, "ref_driver_tracker.cpp"), "-o", os.path.join(OUT, "libdslam_ref_tracker.so")]
        subprocess.run(cmd, check=True)
        print("built", os.path.join(OUT, "libdslam_ref_tracker.so"))
        # the same translation unit with the reference's own optimisation level (CMakeLists.txt:4-11: Release, -march=native;
        # x86-64-v3 instead of native because the library travels to the GPU box) for the timed CPU baseline of bench.py
        cmd_opt = [c for c in cmd if c not in ("-O2", "-ffp-contract=off", "-msse2")]
        cmd_opt[1:1] = ["-O3", "-march=x86-64-v3"]
        cmd_opt[-1] = os.path.join(OUT, "libdslam_ref_tracker_opt.so")
        subprocess.run(cmd_opt, check=True)
        print("built", cmd_opt[-1])
        # ---- PoseEstimator.cpp as a whole (loop-closure alignment) -------------------------------------
        cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-msse2", "-ffp-contract=off", "-w",
               "-I", os.path.join(HERE, "shim"), "-I", tmp, "-I", os.path.join(tmp, "dso", "src"), "-I", os.path.join(REF, "src"),
               "-I", os.path.join(REF, "src", "loop_closure", "pose_estimation"),
               os.path.join(HERE, "ref_driver_pe.cpp"), "-o", os.path.join(OUT, "libdslam_ref_pe.so")]
        subprocess.run(cmd, check=True)
        print("built", os.path.join(OUT, "libdslam_ref_pe.so"))
        # ---- the C++ adapter instantiated with the reference's types, next to the reference's own tracker (adapter_vs_reference) ----
        repo = os.path.dirname(HERE)
        libdir = os.path.join(repo, "direct_stereo_slam_b200")
        if os.path.exists(os.path.join(libdir, "libdslam_b200.so")):
            cmd = ["g++", "-O2", "-std=c++14", "-msse2", "-ffp-contract=off", "-w",
                   "-I", os.path.join(HERE, "shim"), "-I", tmp, "-I", os.path.join(tmp, "dso", "src"),
                   "-I", os.path.join(REF, "src"), "-I", os.path.join(REF, "src", "scale_optimization"), "-I", os.path.join(repo, "include"),
                   os.path.join(HERE, "adapter_vs_reference.cpp"), "-L", libdir, "-ldslam_b200",
                   "-Wl,-rpath,$ORIGIN/../../direct_stereo_slam_b200", "-o", os.path.join(OUT, "adapter_vs_reference")]
            subprocess.run(cmd, check=True)
            print("built", os.path.join(OUT, "adapter_vs_reference"))
        # ---- ScanContext.cpp as a whole (descriptor generation) against oracle/shim_sc ----------------------
        cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-w",
               "-I", os.path.join(HERE, "shim_sc"), "-I", os.path.join(REF, "src"),
               os.path.join(HERE, "ref_driver_sc.cpp"), "-o", os.path.join(OUT, "libdslam_ref_sc.so")]
        subprocess.run(cmd, check=True)
        print("built", os.path.join(OUT, "libdslam_ref_sc.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
