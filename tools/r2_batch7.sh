#!/bin/bash
# round-2 GPU batch 7 (1 GPU): parity-aligned refill, A/B
O=gpurun_out/r2g; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
timeout 600 python -m pytest tests -m gpu -q -x -k "nuts or sampler or every_kernel or parity_build" 2>&1 | tail -8 > $O/gpu_tests.log
for w in "gauss 18" "gauss 20" "PRMwCD 17" "PRMwCD 20"; do
  echo "== noalign $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_noalign.so timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
  echo "== align   $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
done
for w in "arma 20" "arma 17"; do
  echo "== default(noalign) $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
  echo "== alignall $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_alignall.so timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
done
tail -4 $O/gpu_tests.log; cat $O/ab.log
