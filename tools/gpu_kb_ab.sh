#!/bin/bash
# KBRL kernels A/B by build tag: per-phase times at steps 300 and 1000 of the config-3 loop
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  timeout 600 python tools/kbrl_loop.py --envs 16384 --steps 980 --warm 20 --report 300,1000 --resident --dict-cap 2048 2>> gpurun_out/bench.err | python -c "
import json,sys; k=json.loads(sys.stdin.read())
for c in k['checkpoints']: print('$tag step %d: env %.3f  update %.3f  select %.3f  digest %s' % (c['after_steps'], c['ms_env'], c['ms_update_control'], c['ms_select_action'], c['digest_sizes']))
print('$tag overall %.4g env-steps/s' % k['env_steps_per_s'])"
done
