#!/bin/bash
# software-pipelined LSE / normalise kernels: GPU tests, config-5 chain, per-kernel launch list
O=gpurun_out/r2n; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/gpu_tests.log
timeout 600 python bench.py --workload micro --steps 10 --warmup 3 > $O/bench_micro.json 2> $O/bench_micro.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 60 --csv \
      --log-file $O/launches_micro.csv python bench.py --workload micro --steps 3 --warmup 3 > $O/ncu_micro.log 2>&1
cat $O/gpu_tests.log
python - <<'PY'
import json, csv, collections
d = json.loads([l for l in open('gpurun_out/r2n/bench_micro.json') if l.startswith('{')][0])
print(d["ms_per_step"], d["roofline"]["frac"]); print(json.dumps(d["kernels"]))
rows=[r for r in csv.reader(open('gpurun_out/r2n/launches_micro.csv')) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
last={}
for r in rows[1:]:
    if r[ix['Metric Name']]=='gpu__time_duration.sum': last[r[ix['Kernel Name']][:50]]=float(r[ix['Metric Value']].replace(',',''))/1e3
for k,v in last.items(): print(f'{k:50s} {v:8.1f} us')
PY
