#!/bin/bash
# heavy list (units with a long PF loop -> warp-per-unit kernel) under four workloads: list cap x threshold
mkdir -p gpurun_out
for cfg in "0 0" "1000 296" "1000 592" "1000 1184" "1500 592" "2000 1184" "600 296"; do set -- $cfg; thr=$1; cap=$2
  export RS_HEAVY_PF=$thr RS_HEAVY_CAP=$cap
  k=$(timeout 300 python tools/kbrl_loop.py --envs 16384 --steps 280 --warm 20 --report 300 --resident 2>/dev/null | python -c "import json,sys; k=json.loads(sys.stdin.read()); print('%.2f' % k['ms_env'])")
  r=""
  for envs in 4096 16384 65536; do
    v=$(timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f' % d['ms_per_step'])")
    r="$r  random@$envs $v"
  done
  echo "thr $thr cap $cap: kbrl@16384 env $k ms |$r"
done
