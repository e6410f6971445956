#!/bin/bash
O=gpurun_out/r2ag; mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_arma_nvml.json 2> $O/bench_arma_nvml.err
SMCB_BENCH_NVIDIA_SMI=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_arma_smi.json 2> $O/bench_arma_smi.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_arma_nvml2.json 2> $O/bench_arma_nvml2.err
python - <<'PY'
import json
for f in ['nvml','smi','nvml2']:
    d=json.loads([l for l in open(f'gpurun_out/r2ag/bench_arma_{f}.json') if l.startswith('{')][0])
    print(f, 'value %.4g'%d['value'], 'ms', round(d['ms_per_step'],4), 'e2e %.4g'%d['e2e']['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), d['clocks'])
PY
cat $O/*.err | tail -5
