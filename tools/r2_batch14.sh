#!/bin/bash
# occupancy vs shard size: resident CTAs per SM for small shards (the 8-GPU shard of BASELINE configs 2 and 3 is 2^17)
O=gpurun_out/r2m; mkdir -p $O
for k in 1 2 3 4; do
  echo "== arma 16,17,18 at $k CTAs/SM" >> $O/ab.log
  SMCB_NUTS_BLOCKS_PER_SM=$k timeout 300 python tools/ab_time.py arma 16,17,18 5 >> $O/ab.log 2>&1
done
for k in 1 2 3 4; do
  echo "== PRMwCD 17 at $k CTAs/SM" >> $O/ab.log
  SMCB_NUTS_BLOCKS_PER_SM=$k timeout 300 python tools/ab_time.py PRMwCD 17 2 >> $O/ab.log 2>&1
done
cat $O/ab.log
