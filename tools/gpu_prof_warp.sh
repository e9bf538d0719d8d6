#!/bin/bash
# full captures of the warp-per-unit kernel as shipped: whole small batch (2048 envs, automatic route) and the multiplexed L1 (16 384 envs)
mkdir -p gpurun_out
timeout 800 ncu --set full --import-source on --clock-control none -k regex:embb_step_warp --launch-skip 605 --launch-count 1 -f -o gpurun_out/prof_warp_2048 \
  python bench.py --envs-per-gpu 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/prof_warp.log 2>&1; tail -1 gpurun_out/prof_warp.log
timeout 800 ncu --set full --import-source on --clock-control none -k regex:embb_step_warp --launch-skip 205 --launch-count 1 -f -o gpurun_out/prof_mux_warp \
  python tools/mux_bench.py > gpurun_out/prof_mux.log 2>&1; tail -1 gpurun_out/prof_mux.log
