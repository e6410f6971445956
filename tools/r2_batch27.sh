#!/bin/bash
# level-0 U-turn checkpoint loaded before the evaluation: A/B (build -DSMCB_NUTS_PREFETCH_CK=0 vs shipped)
O=gpurun_out/r2ac; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
for w in "arma 16,17,20" "PRMwCD 17,20"; do
  echo "== no prefetch $w" >> $O/ab.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_nopfck.so timeout 300 python tools/ab_time.py $w 5 >> $O/ab.log 2>&1
  echo "== prefetch    $w" >> $O/ab.log; timeout 300 python tools/ab_time.py $w 5 >> $O/ab.log 2>&1
done
timeout 800 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread 2>&1 | tail -5 >> $O/ab.log
cat $O/ab.log
