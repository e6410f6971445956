#!/bin/bash
# round-2 multi-GPU measurement on ONE 8-GPU box:  gpurun --gpus 8 --timeout 1200 -- 'bash tools/r2_scale8.sh'
#   strong scaling (the stated particle counts IN TOTAL) of config 2 (arma, 2^20), config 3 (PRMwCD, 2^20) and
#   config 4 (Gaussian D=100, 2^22, Gaussian-approx L with the moment all-reduce) at 8 GPUs (+ arma at 2, 4),
#   config 5 (micro, 2^25 per GPU, fused peer-store migration), the sharded-equals-unsharded GPU test and check.
set -u
O=gpurun_out/r2s; mkdir -p $O
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 > $O/gpu_multi_test.log
run 2 tools/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1
run 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/strong_arma_n8.json 2> $O/strong_arma_n8.err
run 4 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > $O/strong_arma_n4.json 2> $O/strong_arma_n4.err
run 2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/strong_arma_n2.json 2> $O/strong_arma_n2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/strong_arma_n1.json 2> $O/strong_arma_n1.err
run 8 bench.py --gpus 8 --workload PRMwCD --steps 5 --warmup 3 --no-cpu-baseline > $O/strong_prm_n8.json 2> $O/strong_prm_n8.err
run 8 bench.py --gpus 8 --workload gauss --steps 3 --warmup 3 --no-cpu-baseline > $O/strong_gauss_n8.json 2> $O/strong_gauss_n8.err
run 8 bench.py --gpus 8 --workload micro --steps 10 --warmup 3 > $O/scale_micro_n8.json 2> $O/scale_micro_n8.err
run 8 bench.py --gpus 8 --scaling weak --steps 10 --warmup 5 --no-cpu-baseline > $O/weak_arma_n8.json 2> $O/weak_arma_n8.err
tail -2 $O/gpu_multi_test.log; tail -1 $O/multi_gpu_check.log
for f in strong_arma_n1 strong_arma_n2 strong_arma_n4 strong_arma_n8 strong_prm_n8 strong_gauss_n8 scale_micro_n8 weak_arma_n8; do echo $f; grep '^{' $O/$f.json | cut -c1-230; done
