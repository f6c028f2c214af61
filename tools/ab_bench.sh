#!/bin/bash
# Developer A/B: time the default library and every tools/ab/*.so with the same short bench (ncu launch list per variant).
for lib in default tools/ab/*.so; do
  if [ "$lib" = default ]; then unset FR_LIB_PATH; else export FR_LIB_PATH=$PWD/$lib; fi
  name=$(basename $lib .so)
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-parity > gpurun_out/ab_$name.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file gpurun_out/ab_$name.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > /dev/null 2>&1
done
