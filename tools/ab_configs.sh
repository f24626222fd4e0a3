#!/bin/bash
# A/B of the other configurations on one box: lib/base.so against the current build.
L=mav_trajectory_generation_cmake_b200/lib
cp $L/libminsnap_b200.so $L/new.so
for v in base new base new; do
  cp $L/$v.so $L/libminsnap_b200.so
  echo "== $v"; python tools/bench_configs.py | grep -v shard
done
cp $L/new.so $L/libminsnap_b200.so
