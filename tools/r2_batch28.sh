#!/bin/bash
# other edge loaded before the evaluation at doubling ends: A/B (build -DSMCB_NUTS_PREFETCH_OTHER=0 vs shipped)
O=gpurun_out/r2ad; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
echo "== no other-edge prefetch arma" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_nopfo.so timeout 300 python tools/ab_time.py arma 16,17,20 5 >> $O/ab.log 2>&1
echo "== other-edge prefetch    arma" >> $O/ab.log; timeout 300 python tools/ab_time.py arma 16,17,20 5 >> $O/ab.log 2>&1
echo "== no other-edge prefetch arma" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_nopfo.so timeout 300 python tools/ab_time.py arma 17,20 5 >> $O/ab.log 2>&1
echo "== other-edge prefetch    arma" >> $O/ab.log; timeout 300 python tools/ab_time.py arma 17,20 5 >> $O/ab.log 2>&1
cat $O/ab.log
