#!/bin/bash
# KBRL kernel variants in the resident control loop (config-3 shape).  usage: gpu_kb_sweep.sh tag...
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  echo "== $tag"; python tools/kbrl_loop.py --envs 16384 --steps 30 --warm 170 --dict-cap 128 --resident | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('env_steps_per_s','ms_env','ms_update_control','ms_select_action','digest_sizes','dict_mean')})"
done
