#!/bin/bash
O=gpurun_out/r2ao; mkdir -p $O
timeout 300 python bench.py --workload micro --steps 3 --warmup 3 > $O/plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'lse_partial|normalise_tile_sum' -s 8 -c 2 -o $O/hbm_prof \
    python bench.py --workload micro --steps 3 --warmup 3 > $O/ncu.log 2>&1
ncu -i $O/hbm_prof.ncu-rep --page details --csv > $O/hbm_details.csv 2>/dev/null
ncu -i $O/hbm_prof.ncu-rep --page source --csv > $O/hbm_src.csv 2>/dev/null
ncu -i $O/hbm_prof.ncu-rep --page raw --csv > $O/hbm_raw.csv 2>/dev/null
rm -f $O/hbm_prof.ncu-rep
python - <<'PY'
import csv, collections
rows=list(csv.DictReader(open('gpurun_out/r2ao/hbm_details.csv')))
ker=collections.OrderedDict()
for r in rows: ker.setdefault((r['ID'], r['Kernel Name'][:40]),{})[r['Metric Name']]=r['Metric Value']
for k,d in ker.items(): print(k,{m:d.get(m) for m in ['Duration','Registers Per Thread','DRAM Throughput','Compute (SM) Throughput','Achieved Occupancy','Executed Ipc Active','Issue Slots Busy','Avg. Active Threads Per Warp','Theoretical Occupancy']})
PY
