#!/bin/bash
# round-2 GPU batch 2 (2 GPUs): full GPU suite incl. the sharded test, strong-scaling bench lines, tempering A/B
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > $O/gpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 20 --warmup 5 > $O/bench_arma_n1.json 2> $O/bench_arma_n1.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_arma_n2.json 2> $O/bench_arma_n2.err
python bench.py --workload PRMwCD --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_prm_n1.json 2> $O/bench_prm_n1.err
$TR bench.py --gpus 2 --workload PRMwCD --steps 10 --warmup 3 > $O/bench_prm_n2.json 2> $O/bench_prm_n2.err
python tools/temper_time.py 17 > $O/temper_n1.log 2>&1
$TR tools/temper_time.py 17 > $O/temper_n2.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $O/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> $O/sanitizer.log
tail -5 $O/gpu_tests.log; for f in $O/bench_*.json; do echo $f; cut -c1-400 $f; done; cat $O/temper_*.log; tail -3 $O/sanitizer.log
