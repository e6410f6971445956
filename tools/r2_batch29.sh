#!/bin/bash
O=gpurun_out/r2ae; mkdir -p $O
timeout 300 python tools/ab_step.py 20 25 > $O/ab_step.log 2>&1
timeout 300 python tools/ab_step.py 17 45 >> $O/ab_step.log 2>&1
timeout 800 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread 2>&1 | tail -12 >> $O/ab_step.log
cat $O/ab_step.log
