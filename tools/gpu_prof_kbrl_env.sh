#!/bin/bash
# full captures of the lane-per-unit kernel at 16 384 envs: under the KBRL policy (step ~305 of the control loop, heavy list off)
# and under the random-simplex policy of the bench (after its 600-step burn-in)
mkdir -p gpurun_out
timeout 800 ncu --set full --import-source on --clock-control none -k regex:embb_step_smem --launch-skip 305 --launch-count 1 -f -o gpurun_out/prof_smem_kbrl_16384 \
  python tools/kbrl_loop.py --envs 16384 --steps 10 --warm 300 --resident --heavy 0 > gpurun_out/prof_kbrl.log 2>&1; tail -2 gpurun_out/prof_kbrl.log
timeout 800 ncu --set full --import-source on --clock-control none -k regex:embb_step_smem --launch-skip 605 --launch-count 1 -f -o gpurun_out/prof_smem_random_16384 \
  python bench.py --envs-per-gpu 16384 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/prof_random.log 2>&1; tail -2 gpurun_out/prof_random.log
ls -la gpurun_out/*.ncu-rep
