#!/bin/bash
# Developer A/B on the GPU box: tools/prof_step.py with the default library and with every tools/ab/*.so
for lib in default tools/ab/*.so; do
  if [ "$lib" = default ]; then unset FR_LIB_PATH; else export FR_LIB_PATH=$PWD/$lib; fi
  for B in ${@:-64}; do
    echo "== $lib B=$B"
    python tools/prof_step.py $B 4 2>&1 | grep -v clusters
  done
done
