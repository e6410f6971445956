#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
for k in 1 2 3 4; do
  echo "== noreserve arma 20, $k CTAs/SM" >> $O/ab.log
  SMCB_NUTS_BLOCKS_PER_SM=$k SMCB_LIB_PATH=$L/libsmcnuts_b200_noreserve.so timeout 300 python tools/ab_time.py arma 20 3 >> $O/ab.log 2>&1
done
for k in 1 2 3 4 5 6; do
  echo "== mb6 (80 regs) arma 20, $k CTAs/SM" >> $O/ab.log
  SMCB_NUTS_BLOCKS_PER_SM=$k SMCB_QUEUE=0 SMCB_LIB_PATH=$L/libsmcnuts_b200_mb6.so timeout 300 python tools/ab_time.py arma 20 3 >> $O/ab.log 2>&1
done
SMCB_LIB_PATH=$L/libsmcnuts_b200_noreserve.so timeout 300 python tools/quick_time.py logp >> $O/ab.log 2>&1
cat $O/ab.log
