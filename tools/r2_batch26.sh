#!/bin/bash
O=gpurun_out/r2aa; mkdir -p $O
timeout 300 python tools/ab_step.py 20 25 > $O/ab_step.log 2>&1
timeout 300 python tools/ab_step.py 17 45 >> $O/ab_step.log 2>&1
timeout 800 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread 2>&1 | tail -5 >> $O/ab_step.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_arma.json 2> $O/bench_arma.err
cat $O/ab_step.log; cut -c1-230 $O/bench_arma.json
