#!/bin/bash
# single finish() exit + rolled cold loops: A/B against the previous commit's build, all three NUTS kernels
O=gpurun_out/r2x; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
for w in "gauss 18,20" "arma 17,20" "PRMwCD 17,20"; do
  echo "== prev $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_prev.so timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
  echo "== new  $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
done
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method thread 2>&1 | tail -4 >> $O/ab.log
cat $O/ab.log
