#!/bin/bash
# round-2 GPU batch 5 (1 GPU): Gaussian D=100 NUTS kernel with shared-memory staging of the U-turn operands, A/B
O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests -m gpu -q -x -k "gauss or Gauss or every_kernel or hostsim or nuts" 2>&1 | tail -8 > $O/gpu_tests.log
for lg in 18 20; do
  SMCB_LIB_PATH=$PWD/smc-nuts_b200/smcnuts/_lib/libsmcnuts_b200_nostage.so python tools/ab_time.py gauss $lg 3 >> $O/ab_gauss.log 2>&1
  python tools/ab_time.py gauss $lg 3 >> $O/ab_gauss.log 2>&1
done
python tools/ab_time.py gauss 18 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nuts_transition -s 3 -c 1 -o $O/gauss_prof python tools/ab_time.py gauss 18 2 > $O/ncu_gauss.log 2>&1
tail -4 $O/gpu_tests.log; cat $O/ab_gauss.log
