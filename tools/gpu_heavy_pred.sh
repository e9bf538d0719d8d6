#!/bin/bash
# heavy-list membership from the raw previous count [raw] vs the count scaled to the new allocation [base]: default route at
# 4096 envs (threshold sweep) and the env step under the KBRL policy at 16 384 envs
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  for thr in 300 600 1000 1500; do
  RS_HEAVY_PF=$thr timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu 4096 --no-cpu-baseline --no-configs > gpurun_out/bench_hp_${tag}_$thr.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_hp_${tag}_$thr.json')); print('$tag 4096 envs thr $thr: ms/step %.3f' % d['ms_per_step'])"
  done
  for thr in 600 1000 1500 2500; do
  timeout 300 python tools/kbrl_loop.py --envs 16384 --steps 280 --warm 20 --report 300 --resident --heavy $thr 2>> gpurun_out/bench.err | python -c "
import json,sys; k=json.loads(sys.stdin.read()); print('$tag kbrl@16384 thr $thr step 300: env %.3f ms  total %.3f ms/step  digest %s' % (k['ms_env'], k['ms_per_step_wall'], k['digest_sizes']))"
  done
done
