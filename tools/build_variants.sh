#!/bin/bash
# Builds A/B variants of libdslam_b200.so into scratch/variants/ (git-ignored; travels to the GPU box): the eval kernel with
# other pipeline depths / CTA-per-SM targets, and optionally an older kernels_residual.cu kept under scratch/.
#   usage: tools/build_variants.sh "name:flags[:residual_source]" ...     e.g.  "p3c5:-DDSLAM_EVAL_PIPE=3 -DDSLAM_EVAL_MIN_CTAS=5"
set -e
cd "$(dirname "$0")/../direct_stereo_slam_b200/csrc"
OUT=../../scratch/variants
mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-std=c++17 -O3 $ARCH -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-w"
for spec in "$@"; do
  name="${spec%%:*}"; rest="${spec#*:}"; flags="${rest%%:*}"; src="kernels_residual.cu"
  if [[ "$rest" == *:* ]]; then src="${rest#*:}"; fi
  nvcc $FLAGS $flags -I. -c "$src" -o $OUT/kr_$name.o
  nvcc $ARCH -shared -cudart static -o $OUT/libdslam_b200_$name.so $OUT/kr_$name.o _obj/kernels_pyramid.o _obj/kernels_template.o _obj/kernels_sc.o _obj/dslam_api.o _obj/dslam_sc.o -ldl -lpthread
  echo built $OUT/libdslam_b200_$name.so
done
