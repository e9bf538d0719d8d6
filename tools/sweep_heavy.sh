#!/bin/bash
# default route with the heavy list (units with a long PF loop -> warp-per-unit kernel, concurrent): parity, then the
# threshold sweep at three batch sizes
mkdir -p gpurun_out
RS_HEAVY_PF=150 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "vs_oracle_seeded or golden_B or steady_state or full_size" 2>&1 | tail -5 > gpurun_out/pytest_heavy.txt; cat gpurun_out/pytest_heavy.txt
for envs in 4096 16384 65536; do for thr in 0 300 600 1000 1500 2500; do
RS_HEAVY_PF=$thr timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --no-cpu-baseline --no-configs > gpurun_out/bench_h${thr}_$envs.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_h${thr}_$envs.json')); print('envs $envs heavy_thr $thr: value %.4g  ms/step %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done; done
