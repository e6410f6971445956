#!/bin/bash
# round-2 GPU batch 11 (1 GPU): work-queue reserve + L2 prefetch A/B, 5 / 6 resident CTAs per SM for arma, launch lists of the
# config-5 chain and of a 2^17-particle arma step (the 8-GPU shard size)
O=gpurun_out/r2j; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/gpu_tests.log
for v in noreserve main chunk2 mb5 mb6; do
  echo "== $v arma 17,20" >> $O/ab.log
  if [ $v = main ]; then timeout 300 python tools/ab_time.py arma 17,20 5 >> $O/ab.log 2>&1
  else SMCB_LIB_PATH=$L/libsmcnuts_b200_$v.so timeout 300 python tools/ab_time.py arma 17,20 5 >> $O/ab.log 2>&1; fi
done
for v in noreserve main; do
  echo "== $v PRMwCD scalar 16" >> $O/ab.log
  if [ $v = main ]; then SMCB_PRM_SCALAR=1 timeout 300 python tools/ab_time.py PRMwCD 16 2 >> $O/ab.log 2>&1
  else SMCB_PRM_SCALAR=1 SMCB_LIB_PATH=$L/libsmcnuts_b200_$v.so timeout 300 python tools/ab_time.py PRMwCD 16 2 >> $O/ab.log 2>&1; fi
done
timeout 600 python bench.py --workload micro --steps 3 --warmup 3 > $O/micro_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 60 --csv \
      --log-file $O/launches_micro.csv python bench.py --workload micro --steps 3 --warmup 3 > $O/ncu_micro.log 2>&1
timeout 600 python bench.py --log2n 17 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_arma_n17.json 2> $O/bench_arma_n17.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_arma_n17.csv \
      python bench.py --log2n 17 --steps 5 --warmup 3 --no-cpu-baseline > $O/ncu_arma_n17.log 2>&1
cat $O/gpu_tests.log; cat $O/ab.log; cut -c1-200 $O/bench_arma_n17.json
