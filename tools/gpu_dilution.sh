#!/bin/bash
# lane dilution of the lane-per-unit route (with the heavy list at its per-dilution default threshold) by batch size
mkdir -p gpurun_out
for envs in 4096 6144 8192 12288 16384 24576; do for dil in 0 1 2; do
  RS_WARP_AUTO_UNITS=1 RS_DILUTION=$dil timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --no-cpu-baseline --no-configs > gpurun_out/bench_dil.json 2>> gpurun_out/bench.err
  python -c "
import json; a=json.load(open('gpurun_out/bench_dil.json')); print(json.dumps({'envs': $envs, 'dilution': $dil, 'ms_per_step': round(a['ms_per_step'],3)}))" | tee -a gpurun_out/dilution.jsonl
done; done
