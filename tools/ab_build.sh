#!/bin/bash
# Developer A/B: build library variants into tools/ab/<name>.so;  usage: tools/ab_build.sh name "-DFLAG=1 -DOTHER=2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/ab
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -C 3dfacerecon_b200/csrc -B OUT=$PWD/tools/ab/$name.so EXTRA="$flags" > /dev/null
  grep -A2 "raster_tile_keys_kernel\|recon_fwd_f16_kernelILb1" 3dfacerecon_b200/csrc/ptxas.log | grep -E "registers" | sed "s/^/$name: /"
done
