#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $O/mem.txt
for i in 1 2; do timeout 600 python bench.py --workload micro --steps 10 --warmup 3 > $O/bench_micro_$i.json 2> $O/bench_micro_$i.err; done
python - <<'PY'
import json
for i in (1,2):
    d = json.loads([l for l in open(f'gpurun_out/r2o/bench_micro_{i}.json') if l.startswith('{')][0])
    print(d["ms_per_step"], d["roofline"]["frac"], {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
cat $O/mem.txt
