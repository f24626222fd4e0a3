#!/bin/bash
# A/B on one box: lib/base.so (a build of an earlier commit) against the current build, alternating.
# usage (on the GPU box): bash tools/ab.sh [quick_bench args]
L=mav_trajectory_generation_cmake_b200/lib
cp $L/libminsnap_b200.so $L/new.so
for rep in 1 2 3; do
  for v in base new; do
    cp $L/$v.so $L/libminsnap_b200.so
    echo -n "$v: "; python tools/quick_bench.py --steps 200 "$@" | head -1
  done
done
cp $L/new.so $L/libminsnap_b200.so
