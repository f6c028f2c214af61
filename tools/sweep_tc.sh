#!/bin/bash
# Developer sweep (GPU): recon_project_forward group time of bench.py for tensor-core kernel tuning overrides.
for sc in 1 2 4; do for ge in "2 4" "2 8" "4 4"; do set -- $ge
  r=$(FR_TC_STAGE_CHUNKS=$sc FR_TC_GROUPS=$1 FR_TC_EPI_WARPS=$2 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f us recon, %.1f us step' % (1e3*d['roofline']['groups_ms']['recon_project_forward'], 1e3*d['ms_per_step']))" 2>&1 | tail -1)
  echo "stage_chunks=$sc groups=$1 epi_warps=$2 : $r"
done; done
