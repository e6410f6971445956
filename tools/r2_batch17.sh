#!/bin/bash
# e2e path (NUTSProposal.rvs with pinned host buffers): chunk launches side by side on shares of every SM vs one after the other
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python tools/e2e_time.py > $O/e2e_time.log 2>&1
cat $O/e2e_time.log
