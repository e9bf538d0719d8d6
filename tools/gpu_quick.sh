#!/bin/bash
# Quick GPU check of a kernel change: parity tests + one bench line (no CPU arm, no side configs).  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('value %.4g  ms/step %.3f  e2e %.4g (blocking %.4g)  kernel_ms %.3f  side %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['blocking_value'], d['roofline']['kernel_ms'], d['roofline']['side_kernels_ms']))"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --envs-per-gpu 4096 --no-cpu-baseline --no-configs > gpurun_out/bench_4096.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_4096.json')); print('4096 envs: value %.4g  ms/step %.3f' % (d['value'], d['ms_per_step']))"
