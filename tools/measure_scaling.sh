#!/bin/bash
# Multi-GPU measurement on one 8-GPU box:  gpurun --gpus 8 --timeout 1200 -- 'bash tools/measure_scaling.sh r1b'
#   weak scaling of the default bench (arma, 2^20 particles per GPU) at 8 GPUs,
#   strong scaling of BASELINE config 3 (PRMwCD, 2^20 particles in total) at 2, 4, 8 GPUs,
#   and the sharded-equals-unsharded check at 2 ranks.
set -u
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
case "${2:-arma8}" in *arma8*)
run 2 tools/multi_gpu_check.py > $OUT/multi_gpu_check.log 2>&1; tail -1 $OUT/multi_gpu_check.log
run 8 bench.py --gpus 8 --steps 10 --warmup 5 > $OUT/scale_arma_n8.json 2> $OUT/scale_arma_n8.err; cut -c1-160 $OUT/scale_arma_n8.json
;; esac
WHICH=${2:-"arma8 prm2 prm4 prm8"}
for n in 2 4 8; do
  case "$WHICH" in *prm$n*) ;; *) continue;; esac
  case $n in 2) lg=19;; 4) lg=18;; 8) lg=17;; esac
  run $n bench.py --gpus $n --workload PRMwCD --log2n $lg --steps 4 --warmup 2 --no-cpu-baseline > $OUT/strong_prm_n$n.json 2> $OUT/strong_prm_n$n.err; cut -c1-160 $OUT/strong_prm_n$n.json
done
