#!/bin/bash
# pose solver: group-size sweep at the final configuration (2 CTAs/SM) + DRAM bytes / L2 hit rate of the sweep's launches
set -u
O=gpurun_out/r2_z
mkdir -p $O
timeout 600 python tools/pose_probe.py 32 > $O/pose_probe.txt 2>&1; cat $O/pose_probe.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pose_solve --csv --log-file $O/pose_ncu.csv python tools/pose_probe.py 32 > /dev/null 2>&1
python - <<PY
import csv,io
rows=[l for l in open("$O/pose_ncu.csv") if l.startswith('"')]
d={}
for r in csv.DictReader(io.StringIO("".join(rows))):
    d.setdefault(int(r["ID"]),{"grid":r["Grid Size"]})[r["Metric Name"]]=r["Metric Value"]
for k,v in d.items(): print(k,v)
PY
