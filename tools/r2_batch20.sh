#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
export SMCB_PRM_SCALAR=1
for v in nocg cgnotail main; do
  for args in "6000" "6000 wild" "200 wild"; do
    echo "== $v $args" >> $O/dbg.log
    if [ $v = main ]; then timeout 40 python tools/dbg_prm_scalar.py $args >> $O/dbg.log 2>&1; echo "rc=$?" >> $O/dbg.log
    else SMCB_LIB_PATH=$L/libsmcnuts_b200_$v.so timeout 40 python tools/dbg_prm_scalar.py $args >> $O/dbg.log 2>&1; echo "rc=$?" >> $O/dbg.log; fi
  done
done
cat $O/dbg.log
