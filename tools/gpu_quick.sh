#!/bin/bash
# Quick GPU check of a kernel change: parity tests + one bench line (no CPU arm).  Outputs -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench.err
if [ -n "$KBRL" ]; then
python tools/kbrl_loop.py --envs 16384 --steps 20 --warm 30 --dict-cap 128 --resident > gpurun_out/kbrl_loop_resident_16384.json 2>> gpurun_out/bench.err; cat gpurun_out/kbrl_loop_resident_16384.json; tail -3 gpurun_out/bench.err
fi
