#!/bin/bash
# Round-2 session A: full GPU test suite (all failures listed), one bench line, the KBRL loop at a growing dictionary.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --scenario 3 --no-cpu-baseline > gpurun_out/bench_scn3.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_scn3.json
timeout 900 python tools/kbrl_loop.py --envs 16384 --steps 280 --warm 20 --report 100,200,300 --resident > gpurun_out/kbrl_loop_200.json 2>> gpurun_out/bench.err; cat gpurun_out/kbrl_loop_200.json; tail -3 gpurun_out/bench.err
