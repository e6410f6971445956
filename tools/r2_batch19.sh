#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 420 python -m pytest tests -m gpu -x -v --timeout 100 --timeout-method thread 2>&1 | grep -E "PASSED|FAILED|Timeout|ERROR|passed|failed|::" | tail -60 > $O/gpu_tests.log
tail -25 $O/gpu_tests.log
