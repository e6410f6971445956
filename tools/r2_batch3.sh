#!/bin/bash
# round-2 GPU batch 3 (1 GPU): full GPU suite, arma bench, ncu source-level capture of the arma NUTS kernel
O=gpurun_out/r2c; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/gpu_tests.log
python bench.py --steps 20 --warmup 5 > $O/bench_arma_n1.json 2> $O/bench_arma_n1.err
python tools/ab_time.py arma 20 3 > $O/ab_arma_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nuts_transition -s 3 -c 1 -o $O/arma_prof python tools/ab_time.py arma 20 3 > $O/ncu_arma.log 2>&1
tail -6 $O/gpu_tests.log; cut -c1-300 $O/bench_arma_n1.json; cat $O/ab_arma_plain.log; tail -3 $O/ncu_arma.log
