#!/bin/bash
# A/B of the unroll factor of the ARMA evaluation loop (variant libraries built with build_ext.py --variant uN -DSMCB_ARMA_UNROLL=N)
O=gpurun_out/r2unroll; mkdir -p $O
L=smc-nuts_b200/smcnuts/_lib
for v in "" _u4 _u16 _u24 ""; do
  echo "== unroll ${v:-_u8 (default)}" >> $O/ab.log
  SMCB_LIB_PATH=$L/libsmcnuts_b200$v.so timeout 200 python tools/ab_time.py arma 17,20 7 >> $O/ab.log 2>&1
done
cat $O/ab.log
