#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
timeout 800 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread 2>&1 | tail -30 > $O/gpu_tests.log
cat $O/gpu_tests.log
