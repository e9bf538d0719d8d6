#!/bin/bash
# full GPU suite, then the warp-per-unit route (variant 3) against the lane-per-unit route with its heavy list (RS_WARP_AUTO_UNITS=1) by batch size
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 1200 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
for envs in 2048 3072 4096 6144 8192 16384; do
  timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --variant 3 --no-cpu-baseline --no-configs > gpurun_out/bench_x3_$envs.json 2>> gpurun_out/bench.err
  RS_WARP_AUTO_UNITS=1 timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --no-cpu-baseline --no-configs > gpurun_out/bench_x0_$envs.json 2>> gpurun_out/bench.err
  python -c "
import json; a=json.load(open('gpurun_out/bench_x3_$envs.json')); b=json.load(open('gpurun_out/bench_x0_$envs.json')); print(json.dumps({'envs': $envs, 'warp_per_unit_ms': round(a['ms_per_step'],3), 'lane_per_unit_plus_heavy_list_ms': round(b['ms_per_step'],3)}))" | tee -a gpurun_out/crossover2.jsonl
done
timeout 300 python tools/mux_bench.py 2>> gpurun_out/bench.err | tee gpurun_out/mux_bench.txt
