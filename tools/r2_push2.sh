#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/push_time.py 25 > $O/push_time_n2.log 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$" $O/push_time_n2.log | tail -12
