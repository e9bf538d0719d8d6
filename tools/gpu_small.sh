#!/bin/bash
# Small-batch behaviour (BASELINE configs[1] = 4096 envs, 16384 envs): latency variant (RS_WIDE) x lane dilution (RS_DILUTION)
for e in 4096 16384; do for w in 0 1; do for d in 0 1 2; do
  RS_WIDE=$w RS_DILUTION=$d python bench.py --steps 20 --warmup 5 --envs-per-gpu $e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('envs $e wide $w dil $d: %.3fM env-steps/s  %.3f ms/step' % (d['value']/1e6, d['ms_per_step']))"
done; done; done
