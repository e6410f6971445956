#!/bin/bash
# Round-end measurement on ONE B200 (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/measure_round.sh r1b'
# Writes gpurun_out/<tag>/: GPU test log, bench lines (arma = default, reference arm, PRMwCD, gauss, micro), the ncu
# launch list of the default bench and --set full captures of the two headline NUTS kernels.  Every ncu command runs
# only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -1 $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $OUT/bench_arma.json 2> $OUT/bench_arma.err && cut -c1-220 $OUT/bench_arma.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_arma_reference.json 2>> $OUT/bench_arma.err
python bench.py --workload PRMwCD --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_prm.json 2> $OUT/bench_prm.err && cut -c1-200 $OUT/bench_prm.json
python bench.py --workload gauss --log2n 20 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_gauss.json 2> $OUT/bench_gauss.err && cut -c1-200 $OUT/bench_gauss.json
python bench.py --workload micro --steps 10 --warmup 3 > $OUT/bench_micro.json 2> $OUT/bench_micro.err && cut -c1-200 $OUT/bench_micro.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/b_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench_arma.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:nuts_transition -s 5 -c 1 -o $OUT/prof_nuts_arma \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_arma.log 2>&1
python tools/ab_time.py PRMwCD 18 1 > $OUT/ab_prm.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:nuts_transition -s 2 -c 1 -o $OUT/prof_nuts_prm \
      python tools/ab_time.py PRMwCD 18 1 > $OUT/ncu_full_prm.log 2>&1
ls -la $OUT
