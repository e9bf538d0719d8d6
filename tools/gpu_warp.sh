#!/bin/bash
# warp-per-unit kernel: parity (variant 3 in the parametrised tests) and the batch-size crossover against the shared-memory kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "golden_B or vs_oracle_seeded" 2>&1 | tail -8 > gpurun_out/pytest_warp.txt; cat gpurun_out/pytest_warp.txt
for envs in 2048 4096 8192 16384 32768 65536; do for v in 3 4; do
timeout 300 python bench.py --steps 20 --warmup 5 --envs-per-gpu $envs --variant $v --no-cpu-baseline --no-configs > gpurun_out/bench_v${v}_$envs.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v${v}_$envs.json')); print('envs $envs variant $v: value %.4g  ms/step %.3f  kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
done; done
