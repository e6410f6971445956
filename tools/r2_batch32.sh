#!/bin/bash
O=gpurun_out/r2ak; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
export SMCB_PRM_SCALAR=1
for v in dbg1 dbg2 dbg3; do
  echo "== $v" >> $O/dbg.log
  SMCB_LIB_PATH=$L/libsmcnuts_b200_$v.so timeout 20 python tools/dbg_prm_scalar.py 200 wild >> $O/dbg.log 2>&1; echo "rc=$?" >> $O/dbg.log
done
cat $O/dbg.log
