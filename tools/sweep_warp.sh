#!/bin/bash
# occupancy variants of the warp-per-unit kernel at small batches: bash tools/sweep_warp.sh base wmb3 wmb4
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  for envs in 2048 4096 8192; do
  timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --variant 3 --no-cpu-baseline --no-configs > gpurun_out/bench_w_${tag}_$envs.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_w_${tag}_$envs.json')); print('$tag envs $envs: value %.4g  ms/step %.3f  kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
  done
done
