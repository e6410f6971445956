#!/bin/bash
# round-2 GPU batch 4 (1 GPU): ncu source-level captures of the Gaussian D=100 and PRMwCD group NUTS kernels
O=gpurun_out/r2d; mkdir -p $O
python tools/ab_time.py gauss 18 2 > $O/ab_gauss_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nuts_transition -s 3 -c 1 -o $O/gauss_prof python tools/ab_time.py gauss 18 2 > $O/ncu_gauss.log 2>&1
python tools/ab_time.py PRMwCD 17 2 > $O/ab_prm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nuts_transition -s 3 -c 1 -o $O/prm_prof python tools/ab_time.py PRMwCD 17 2 > $O/ncu_prm.log 2>&1
python tools/ab_time.py gauss 20 2 >> $O/ab_gauss_plain.log 2>&1
cat $O/ab_gauss_plain.log $O/ab_prm_plain.log; tail -2 $O/ncu_gauss.log $O/ncu_prm.log
