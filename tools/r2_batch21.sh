#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
SMCB_PRM_SCALAR=1 timeout 60 python tools/dbg_prm_scalar.py 6000 wild > $O/dbg.log 2>&1; echo "rc=$?" >> $O/dbg.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method thread 2>&1 | tail -6 > $O/gpu_tests.log
timeout 300 python tools/ab_time.py gauss 18,20 3 > $O/ab.log 2>&1
timeout 600 python bench.py --workload gauss --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_gauss_n22.json 2> $O/bench_gauss_n22.err
cat $O/dbg.log $O/gpu_tests.log $O/ab.log; cut -c1-250 $O/bench_gauss_n22.json
