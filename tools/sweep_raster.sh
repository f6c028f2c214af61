#!/bin/bash
# Developer sweep (GPU): render_depth_forward group time of bench.py for rasterizer tuning overrides.
for fpt in 4 8; do for minb in 1 4; do
  r=$(FR_RASTER_FPT=$fpt FR_RASTER_MINB=$minb timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f us render, %.1f us step' % (1e3*d['roofline']['groups_ms']['render_depth_forward'], 1e3*d['ms_per_step']))" 2>&1 | tail -1)
  echo "fpt=$fpt minblocks=$minb : $r"
done; done
