#!/bin/bash
# round-2 GPU batch 8 (1 GPU): FP64 latency microbenchmark; partial-sum dot products in the wide group kernel; fused
# normalise+scan / vectorised LSE in the HBM chain
O=gpurun_out/r2h; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
tools/bin/fp64_latency > $O/fp64_latency.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/gpu_tests.log
for w in "gauss 18" "gauss 20"; do
  echo "== before (noalign build of the previous commit) $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_noalign.so timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
  echo "== partial sums $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
done
timeout 600 python bench.py --workload micro --steps 10 --warmup 3 > $O/bench_micro.json 2> $O/bench_micro.err
cat $O/fp64_latency.log; tail -3 $O/gpu_tests.log; cat $O/ab.log; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2h/bench_micro.json') if l.startswith('{')][0])
print(d["ms_per_step"], d["roofline"]["frac"]); print(json.dumps(d["kernels"], indent=1))
PY
