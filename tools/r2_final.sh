#!/bin/bash
# final measurement of round 2 on one B200: GPU tests, bench lines of every workload, reference arm, launch list and a
# --set full capture of the headline kernel with the final build
O=gpurun_out/r2final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread 2>&1 | tail -4 > $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_arma.json 2> $O/bench_arma.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_arma_reference.json 2> $O/bench_arma_reference.err
timeout 600 python bench.py --workload PRMwCD --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_prm.json 2> $O/bench_prm.err
timeout 900 python bench.py --workload gauss --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_gauss_n22.json 2> $O/bench_gauss_n22.err
timeout 600 python bench.py --workload micro --steps 10 --warmup 3 > $O/bench_micro.json 2> $O/bench_micro.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/b_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_arma.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 300 python tools/ab_time.py arma 20 1 > $O/ab_arma_plain.log 2>&1 && \
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:nuts_transition -s 2 -c 1 -o $O/arma_prof \
      python tools/ab_time.py arma 20 1 > $O/ncu_arma.log 2>&1
[ -f $O/arma_prof.ncu-rep ] && ncu -i $O/arma_prof.ncu-rep --page details --csv > $O/arma_details.csv 2>/dev/null
[ -f $O/arma_prof.ncu-rep ] && ncu -i $O/arma_prof.ncu-rep --page source --csv > $O/arma_src.csv 2>/dev/null
[ -f $O/arma_prof.ncu-rep ] && ncu -i $O/arma_prof.ncu-rep --page raw --csv > $O/arma_raw.csv 2>/dev/null
rm -f $O/arma_prof.ncu-rep
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
cat $O/pytest_gpu.log
for f in arma arma_reference prm gauss_n22 micro; do grep '^{' $O/bench_$f.json | cut -c1-200; done
