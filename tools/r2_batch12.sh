#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
for v in noreserve nopf chunk8 chunk8nopf; do
  echo "== $v arma 20" >> $O/ab.log
  SMCB_LIB_PATH=$L/libsmcnuts_b200_$v.so timeout 300 python tools/ab_time.py arma 20 5 >> $O/ab.log 2>&1
done
cat $O/ab.log
