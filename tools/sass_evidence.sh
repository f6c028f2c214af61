#!/bin/bash
# Developer tool: SASS mnemonic counts that prove tcgen05 / TMA / DPX / PDL in the shipped library -> profiles/rNN_sass_evidence.txt
LIB=${1:-/root/repo/3dfacerecon_b200/lib3dfacerecon_b200.so}
S=$(cuobjdump -sass "$LIB")
echo "# cuobjdump -sass 3dfacerecon_b200/lib3dfacerecon_b200.so: instruction counts (whole-word matches)"
for m in UTCHMMA LDTM UBLKCP UTCBAR REDG.E.MAX.64 VIMNMX3.U16x2 VIMNMX3 ACQBULK PREEXIT SYNCS HMMA HGMMA; do
  printf "%-16s %s\n" "$m" "$(echo "$S" | grep -cw -- "$m")"
done
echo "(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, VIMNMX3 = 3-input DPX min/max,"
echo " ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents; no legacy HMMA / HGMMA)"
echo "(per-kernel table: see the python snippet in profiles/r02_summary.md)"
