#!/bin/bash
# D = 100 kernel: workspace records through L2 only (ld/st.global.cg) vs through L1
O=gpurun_out/r2r; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
for v in nocg main; do
  echo "== $v gauss 18,20" >> $O/ab.log
  if [ $v = main ]; then timeout 300 python tools/ab_time.py gauss 18,20 3 >> $O/ab.log 2>&1
  else SMCB_LIB_PATH=$L/libsmcnuts_b200_$v.so timeout 300 python tools/ab_time.py gauss 18,20 3 >> $O/ab.log 2>&1; fi
done
timeout 900 python -m pytest tests -m gpu -q -k "gauss or Gauss or parity or kernel_family" 2>&1 | tail -4 >> $O/ab.log
cat $O/ab.log
