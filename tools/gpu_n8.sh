#!/bin/bash
# 8-GPU e2e scaling experiment: per-rank device / blocking / pipelined times with and without CPU binding.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))" >> gpurun_out/n8_topo.txt; lscpu | grep -E "NUMA|Model name|^CPU\(s\)" >> gpurun_out/n8_topo.txt
for mode in none auto; do
  RS_BENCH_AFFINITY=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_bench_$mode.json 2> gpurun_out/n8_bench_$mode.err
  grep "bench rank" gpurun_out/n8_bench_$mode.err | sort; python -c "
import json; d=json.load(open('gpurun_out/n8_bench_$mode.json')); print('$mode value %.4g e2e %.4g pipelined %.4g blocking %.4g' % (d['value'], d['e2e']['value'], d['e2e']['pipelined_value'], d['e2e']['blocking_value']))"
done
