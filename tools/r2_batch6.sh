#!/bin/bash
# round-2 GPU batch 6 (1 GPU): record-header / lazy-start-row / slim RNG changes: tests + A/B against the previous build
O=gpurun_out/r2f; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/gpu_tests.log
OLD=$PWD/smc-nuts_b200/smcnuts/_lib/libsmcnuts_b200_nostage.so
for w in "arma 20" "PRMwCD 17" "PRMwCD 20" "gauss 18" "gauss 20"; do
  echo "== old $w" >> $O/ab.log; SMCB_LIB_PATH=$OLD python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
  echo "== new $w" >> $O/ab.log; python tools/ab_time.py $w 3 >> $O/ab.log 2>&1
done
tail -4 $O/gpu_tests.log; cat $O/ab.log
