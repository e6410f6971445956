#!/bin/bash
# round-2 GPU batch 10 (1 GPU): tail compaction A/B (main build vs -DSMCB_TAIL_COMPACT=0), GPU tests incl. the generated
# Stan-subset models (nvcc on the box) and the step-size adaptation
O=gpurun_out/r2i; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
which nvcc > $O/nvcc.log 2>&1; nvcc --version >> $O/nvcc.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/gpu_tests.log
for w in "arma 17,20" "PRMwCD 17,20" "gauss 8"; do
  echo "== notail $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_notail.so timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
  echo "== tail   $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_arma.json 2> $O/bench_arma.err
SMCB_LIB_PATH=$L/libsmcnuts_b200_notail.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_arma_notail.json 2> $O/bench_arma_notail.err
cat $O/nvcc.log; cat $O/gpu_tests.log; cat $O/ab.log; cut -c1-200 $O/bench_arma.json $O/bench_arma_notail.json
