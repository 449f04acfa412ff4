#!/bin/bash
# A/B build of the library with extra compile flags: bash tools/build_variant.sh TAG "-DFOO=1 ..." [file.cu ...]
# -> gpurun_ab_TAG.so at the repo root (git-ignored, travels to the GPU box); MSCKF_B200_LIB=$PWD/gpurun_ab_TAG.so selects it.
# Only the named sources are recompiled with the flags (default: all); the other objects come from the in-tree build.
set -e
TAG=$1; FLAGS=$2; shift 2 || true
CSRC=monocular_visual_inertial_msckf_b200/csrc
make -s >/dev/null
OBJ=/tmp/msckf_variant_$TAG; rm -rf $OBJ; mkdir -p $OBJ
FILES="$@"; [ -z "$FILES" ] && FILES=$(cd $CSRC && ls *.cu)
LINK=""
for f in $(cd $CSRC && ls *.cu); do
  o=${f%.cu}.o
  if echo " $FILES " | grep -q " $f "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $FLAGS -c $CSRC/$f -o $OBJ/$o 2> $OBJ/$o.log &
    LINK="$LINK $OBJ/$o"
  else
    LINK="$LINK $CSRC/$o"
  fi
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gpurun_ab_$TAG.so $LINK
echo built gpurun_ab_$TAG.so
