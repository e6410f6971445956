#!/bin/bash
# strong-scaling lines at 4 GPUs with the final build (configs 2, 3, 4)
O=gpurun_out/r2s4; mkdir -p $O
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
run 4 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > $O/strong_arma_n4.json 2> $O/strong_arma_n4.err
run 4 bench.py --gpus 4 --workload PRMwCD --steps 5 --warmup 3 --no-cpu-baseline > $O/strong_prm_n4.json 2> $O/strong_prm_n4.err
run 4 bench.py --gpus 4 --workload gauss --steps 3 --warmup 3 --no-cpu-baseline > $O/strong_gauss_n4.json 2> $O/strong_gauss_n4.err
run 2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/strong_arma_n2.json 2> $O/strong_arma_n2.err
for f in strong_arma_n4 strong_prm_n4 strong_gauss_n4 strong_arma_n2; do echo $f; grep '^{' $O/$f.json | cut -c1-230; done
