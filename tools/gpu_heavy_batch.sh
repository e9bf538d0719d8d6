#!/bin/bash
# batched PF step in the heavy-list instantiation [h5] against the shipped library [base], with the heavy threshold re-swept
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  for thr in 500 700 1000; do
  RS_HEAVY_PF=$thr timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu 4096 --no-cpu-baseline --no-configs > gpurun_out/bench_hb.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_hb.json')); print('$tag 4096 envs thr $thr: ms/step %.3f' % d['ms_per_step'])"
  done
  for thr in 1000 1500; do
  RS_HEAVY_PF=$thr timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu 16384 --no-cpu-baseline --no-configs > gpurun_out/bench_hb.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_hb.json')); print('$tag 16384 envs thr $thr: ms/step %.3f' % d['ms_per_step'])"
  done
  for thr in 1200 1800 2500; do
  timeout 300 python tools/kbrl_loop.py --envs 16384 --steps 280 --warm 20 --report 300 --resident --heavy $thr 2>> gpurun_out/bench.err | python -c "
import json,sys; k=json.loads(sys.stdin.read()); print('$tag kbrl@16384 thr $thr step 300: env %.3f ms  total %.3f ms/step' % (k['ms_env'], k['ms_per_step_wall']))"
  done
done
