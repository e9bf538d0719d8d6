#!/bin/bash
# A/B of warp-kernel builds: parity of the default build, then variant 3 at small batches, the default route at 4096 envs,
# the multiplexed L1 and the env step under the KBRL policy, per build tag (base = the default library)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  for envs in 2048 4096; do
  timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --variant 3 --no-cpu-baseline --no-configs > gpurun_out/bench_w_${tag}_$envs.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_w_${tag}_$envs.json')); print('$tag variant 3 envs $envs: ms/step %.3f  kernel_ms %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
  done
  timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu 4096 --no-cpu-baseline --no-configs > gpurun_out/bench_d_${tag}_4096.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_d_${tag}_4096.json')); print('$tag default route envs 4096: ms/step %.3f' % d['ms_per_step'])"
  timeout 300 python tools/mux_bench.py 2>> gpurun_out/bench.err | sed "s/^/$tag /"
  timeout 300 python tools/kbrl_loop.py --envs 16384 --steps 280 --warm 20 --report 300 --resident 2>> gpurun_out/bench.err | python -c "
import json,sys; k=json.loads(sys.stdin.read()); print('$tag kbrl@16384 step 300: env %.3f ms  update %.3f  select %.3f  digest %s' % (k['ms_env'], k['ms_update_control'], k['ms_select_action'], k['digest_sizes']))"
done
