#!/bin/bash
# Evidence call for the current build: step profile at the bench configuration, ncu launch list with DRAM traffic per launch.
set -u
O=gpurun_out/${1:-s12}
mkdir -p $O
timeout 300 python tools/step_profile.py --pairs 64 --chunk 32 --top 60 > $O/step_profile_pairs64_chunk32.txt 2>&1
# launch list of one bench step (cold-cache, serialised: compare shares, not absolutes); cache control off so L2 state is the run's own
timeout 800 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
  --csv --log-file $O/launches_pairs32.csv python bench.py --steps 1 --warmup 1 --pairs 32 --no-cpu-baseline --latency-frames 0 > $O/ncu_bench.log 2>&1
python - <<PY
import csv, collections, re
rows=[r for r in csv.reader(open("$O/launches_pairs32.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.defaultdict(lambda:[0,0.0,0.0,0.0])
for r in rows[1:]:
    try: v=float(r[ix["Metric Value"]].replace(",",""))
    except ValueError: continue
    name=re.sub(r"\(.*","",r[ix["Kernel Name"]])[:60]; m=r[ix["Metric Name"]]; u=r[ix["Metric Unit"]]
    a=agg[name]
    if m=="gpu__time_duration.sum": a[0]+=1; a[1]+=v*{"ns":1e-3,"us":1.0,"ms":1e3,"s":1e6}.get(u,1.0)
    else:
        b=v*{"byte":1.0,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1.0)
        if m=="dram__bytes_read.sum": a[2]+=b
        else: a[3]+=b
tot=sum(a[1] for a in agg.values())
print("kernel, launches, total_us, share, avg_us, dram_read_MB_per_launch, dram_write_MB_per_launch")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print(f"{k}, {a[0]}, {a[1]:.0f}, {a[1]/tot:.3f}, {a[1]/max(a[0],1):.1f}, {a[2]/max(a[0],1)/1e6:.1f}, {a[3]/max(a[0],1)/1e6:.1f}")
PY
gzip -f $O/launches_pairs32.csv
head -40 $O/step_profile_pairs64_chunk32.txt
