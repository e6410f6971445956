#!/bin/bash
O=gpurun_out/r2aq; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
for v in gram32 main gram128; do
  echo "== $v" >> $O/t.log
  if [ $v = main ]; then timeout 120 python tools/gaussl_time.py 20 >> $O/t.log 2>&1; else SMCB_LIB_PATH=$L/libsmcnuts_b200_$v.so timeout 120 python tools/gaussl_time.py 20 >> $O/t.log 2>&1; fi
done
timeout 300 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread -k "gauss or Gauss or lkernel" 2>&1 | tail -3 >> $O/t.log
cat $O/t.log
