#!/bin/bash
# threshold sweep of the heavy list with the scaled prediction: random policy at 4096 / 16384 / 65536 envs, KBRL policy at 16384
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "vs_oracle_seeded or golden_B or steady_state or automatic_route or every_route" 2>&1 | tail -3
for cfg in "4096 800" "4096 1000" "4096 1200" "16384 0" "16384 1500" "16384 2500" "16384 3500" "65536 0" "65536 3000" "65536 4000"; do set -- $cfg
  RS_HEAVY_PF=$2 timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $1 --no-cpu-baseline --no-configs > gpurun_out/bench_hp2_$1_$2.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_hp2_$1_$2.json')); print('random $1 envs thr $2: ms/step %.3f' % d['ms_per_step'])"
done
for thr in 2500 3500 5000; do
  timeout 300 python tools/kbrl_loop.py --envs 16384 --steps 280 --warm 20 --report 300 --resident --heavy $thr 2>> gpurun_out/bench.err | python -c "
import json,sys; k=json.loads(sys.stdin.read()); print('kbrl@16384 thr $thr step 300: env %.3f ms  total %.3f ms/step  digest %s' % (k['ms_env'], k['ms_per_step_wall'], k['digest_sizes']))"
done
timeout 300 python tools/kbrl_env_profile.py --envs 16384 --steps 300 --thresholds 2500,3500 2>> gpurun_out/bench.err
