#!/bin/bash
# the driver's N = 2 launch of both arms with the final tree
O=gpurun_out/r2check2; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $O/bench_reference_n2.json 2> $O/bench_reference_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
for f in reference_n2 n2; do grep '^{' $O/bench_$f.json | cut -c1-600; tail -2 $O/bench_$f.err; done
