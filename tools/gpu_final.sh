#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_scn0_65536.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_scn0_65536.json')); print(d['value'], d['e2e']['value'], d['e2e']['blocking_value'], d['roofline']['frac'], d.get('issue_roofline',{}).get('frac'))
for c in d['configs']: print(c['config'][:40], c['value'], c.get('ms_per_step'), c.get('ms_env'))"; tail -2 gpurun_out/bench.err
