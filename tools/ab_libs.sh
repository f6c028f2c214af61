#!/bin/bash
# Developer tool: the fused step (tools/prof_step.py, L2 flushed per call) with every library variant under tools/ab/ and
# the in-tree build.   bash tools/ab_libs.sh [B] [reps] > gpurun_out/ab.log
B=${1:-64}; R=${2:-40}
echo "== default"; python tools/prof_step.py $B $R grid tiles | grep fused
for f in tools/ab/*.so; do
  case $f in *lib_tl*) continue;; esac
  echo "== $f"; FR_LIB_PATH=$PWD/$f python tools/prof_step.py $B $R grid tiles | grep fused
done
