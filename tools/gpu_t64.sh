#!/bin/bash
# block size of the lane-per-unit kernel: tags as built with RS_BUILD_TAG (base = default library)
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  if [ "$tag" != base ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "vs_oracle_seeded or golden_B or steady_state or every_route or guard" 2>&1 | tail -2; fi
  for envs in 4096 16384 65536; do
  timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --no-cpu-baseline --no-configs > gpurun_out/bench_t_${tag}_$envs.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_t_${tag}_$envs.json')); print('$tag envs $envs: ms/step %.3f  kernel_ms %.3f  value %.4g' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']))"
  done
  timeout 300 python tools/kbrl_loop.py --envs 16384 --steps 280 --warm 20 --report 300 --resident 2>> gpurun_out/bench.err | python -c "
import json,sys; k=json.loads(sys.stdin.read()); print('$tag kbrl@16384 step 300: env %.3f ms  total %.3f' % (k['ms_env'], k['ms_per_step_wall']))"
done
