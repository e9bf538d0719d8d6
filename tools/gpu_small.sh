#!/bin/bash
# Small-batch behaviour: BASELINE configs[1] (4096 envs), 16384 and 65536 envs, for the given library tags ("base" = default)
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  for e in 4096 16384 65536; do
    python bench.py --steps 20 --warmup 5 --envs-per-gpu $e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag envs $e: %.3fM env-steps/s  %.3f ms/step' % (d['value']/1e6, d['ms_per_step']))"
  done
done
