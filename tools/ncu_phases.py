"""Developer tool: per-phase instruction / stall-sample / shared-wavefront totals of one kernel from an .ncu-rep captured with
--import-source on.  Phases are split at the BAR.SYNC instructions (and at extra SASS-offset marks given on the command line).

    python tools/ncu_phases.py report.ncu-rep [hex offset ...]
"""
import csv, subprocess, sys
rep = sys.argv[1]
marks = [int(x, 16) for x in sys.argv[2:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
base = None
phases, cur = [], {"start": 0, "inst": 0, "samples": 0, "wave": 0, "thr": 0, "n": 0}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[col["Address"]], 16)
    if base is None:
        base = addr
    off = addr - base
    src = r[col["Source"]].strip()
    if off in marks:
        phases.append(cur)
        cur = {"start": off, "inst": 0, "samples": 0, "wave": 0, "thr": 0, "n": 0}
    cur["inst"] += int(r[col["Instructions Executed"]])
    cur["thr"] += int(r[col["Thread Instructions Executed"]])
    cur["samples"] += int(r[col["# Samples"]])
    cur["wave"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    cur["n"] += 1
    if src.startswith("BAR.SYNC"):
        phases.append(cur)
        cur = {"start": off + 16, "inst": 0, "samples": 0, "wave": 0, "thr": 0, "n": 0}
phases.append(cur)
ti = sum(p["inst"] for p in phases) or 1
ts = sum(p["samples"] for p in phases) or 1
for p in phases:
    print("from 0x%04x  %4d SASS  inst %9d (%4.1f%%)  threads/inst %4.1f  samples %6d (%4.1f%%)  smem wavefronts %8d" %
          (p["start"], p["n"], p["inst"], 100.0 * p["inst"] / ti, p["thr"] / max(p["inst"], 1), p["samples"], 100.0 * p["samples"] / ts, p["wave"]))
