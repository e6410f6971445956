#!/bin/bash
O=gpurun_out/r2an; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
for rep in 1 2; do
for w in "arma 17,20" "PRMwCD 20"; do
  echo "== no row prefetch $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_nopfr.so timeout 300 python tools/ab_time.py $w 5 >> $O/ab.log 2>&1
  echo "== row prefetch    $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 5 >> $O/ab.log 2>&1
done
done
cat $O/ab.log
