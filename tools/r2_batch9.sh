#!/bin/bash
# round-2 GPU batch 9 (1 GPU): everything batch 8 wanted (the pod was busy), the step-size adaptation tests, bench lines of
# every workload incl. config 4 at its stated N = 2^22, fresh --set full captures of the arma / Gaussian NUTS kernels and
# of the Gaussian-L / moment kernels.  Every ncu command runs after the same command exited 0 without ncu.
O=gpurun_out/r2h; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
tools/bin/fp64_latency > $O/fp64_latency.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/gpu_tests.log
for w in "gauss 18" "gauss 20"; do
  echo "== before (noalign build of commit 31ba8b2) $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_noalign.so timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
  echo "== partial sums $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
done
echo "== arma 20" >> $O/ab.log; timeout 300 python tools/ab_time.py arma 20 3 >> $O/ab.log 2>&1
echo "== PRMwCD 20" >> $O/ab.log; timeout 300 python tools/ab_time.py PRMwCD 20 2 >> $O/ab.log 2>&1
timeout 600 python bench.py --workload micro --steps 10 --warmup 3 > $O/bench_micro.json 2> $O/bench_micro.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_arma.json 2> $O/bench_arma.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_arma_reference.json 2> $O/bench_arma_reference.err
timeout 600 python bench.py --workload PRMwCD --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_prm.json 2> $O/bench_prm.err
timeout 900 python bench.py --workload gauss --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_gauss_n22.json 2> $O/bench_gauss_n22.err
# ncu: arma and Gaussian NUTS kernels on the fixed inputs of tools/ab_time.py (plain run first)
timeout 300 python tools/ab_time.py arma 20 1 > $O/ab_arma_plain.log 2>&1 && \
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:nuts_transition -s 2 -c 1 -o $O/arma_prof \
      python tools/ab_time.py arma 20 1 > $O/ncu_arma.log 2>&1
timeout 300 python tools/ab_time.py gauss 18 1 > $O/ab_gauss_plain.log 2>&1 && \
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:nuts_transition -s 2 -c 1 -o $O/gauss_prof \
      python tools/ab_time.py gauss 18 1 > $O/ncu_gauss.log 2>&1
# ncu: Gaussian-L kernels and the weighted-moment kernel inside the config-4 bench at N = 2^20
timeout 600 python bench.py --workload gauss --log2n 20 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_gauss_n20.json 2> $O/bench_gauss_n20.err && \
  timeout 900 ncu --set full --clock-control none -k regex:'gaussL|weighted_moment' -s 12 -c 6 -o $O/gaussL_prof \
      python bench.py --workload gauss --log2n 20 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_gaussL.log 2>&1
# launch list of the default bench
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/b_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_arma.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
for f in arma gauss gaussL; do
  [ -f $O/${f}_prof.ncu-rep ] && ncu -i $O/${f}_prof.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
  [ -f $O/${f}_prof.ncu-rep ] && ncu -i $O/${f}_prof.ncu-rep --page details --csv > $O/${f}_details.csv 2>/dev/null
done
for f in arma gauss; do [ -f $O/${f}_prof.ncu-rep ] && ncu -i $O/${f}_prof.ncu-rep --page source --csv > $O/${f}_src.csv 2>/dev/null; done
rm -f $O/gaussL_prof.ncu-rep
cat $O/fp64_latency.log; cat $O/gpu_tests.log; cat $O/ab.log
for f in arma prm gauss_n22 micro; do cut -c1-260 $O/bench_$f.json; done
ls -la $O
