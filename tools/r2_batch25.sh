#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
export SMCB_PRM_SCALAR=1
SMCB_LIB_PATH=$L/libsmcnuts_b200_dbgcg.so timeout 60 python tools/dbg_prm_scalar.py 200 wild > $O/dbg.log 2>&1; echo "rc=$?" >> $O/dbg.log
sort $O/dbg.log | uniq -c | sort -rn | head -30
