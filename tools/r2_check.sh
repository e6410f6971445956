#!/bin/bash
# last check of round 2 on one B200: the driver's own sequence (GPU tests, smoke, reference arm, default bench line)
O=gpurun_out/r2check; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread 2>&1 | tail -4 > $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
cat $O/pytest_gpu.log
for f in reference default; do grep '^{' $O/bench_$f.json | cut -c1-400; done
