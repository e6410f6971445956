#!/bin/bash
# streamed host path (one gated launch): correctness + timing against the chunked pipeline, and A/B of the ungated kernels
O=gpurun_out/r2ai; mkdir -p $O
L=$PWD/smc-nuts_b200/smcnuts/_lib
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 --timeout-method thread -k "rvs_host_pipeline" 2>&1 | tail -3 > $O/t.log
echo "== streamed" >> $O/t.log; timeout 120 python tools/e2e_diag.py 2>&1 | tail -3 >> $O/t.log
echo "== chunked pipeline" >> $O/t.log; SMCB_RVS_STREAMED=0 timeout 120 python tools/e2e_diag.py 2>&1 | tail -3 >> $O/t.log
for w in "arma 17,20" "PRMwCD 20" "gauss 18"; do
  echo "== prev $w" >> $O/t.log; SMCB_LIB_PATH=$L/libsmcnuts_b200_prev.so timeout 300 python tools/ab_time.py $w 3 >> $O/t.log 2>&1
  echo "== new  $w" >> $O/t.log; timeout 300 python tools/ab_time.py $w 3 >> $O/t.log 2>&1
done
cat $O/t.log
