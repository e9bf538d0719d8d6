#!/bin/bash
# ncu evidence for the current build: launch list of the bench command + one full capture of the dominant kernel.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 2420 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --burn-in 600 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:embb_step_smem -s 600 -c 1 -o gpurun_out/prof_smem_65536 -f python tools/ncu_step.py --envs 65536 --burn-in 600 --steps 2 > /dev/null 2>&1
ls -la gpurun_out | tail -5
