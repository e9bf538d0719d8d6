#!/bin/bash
# ncu evidence for K2 (mMTC, scenario_3) and K3 (KBRL kernels in the resident control loop).  Outputs -> gpurun_out/
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 4240 -c 80 --csv --log-file gpurun_out/launches_scn3.csv python bench.py --scenario 3 --steps 4 --warmup 3 --burn-in 600 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mmtc_step_kernel -s 600 -c 1 -o gpurun_out/prof_mmtc_65536 -f python tools/ncu_step.py --scenario 3 --envs 65536 --burn-in 600 --steps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 150 -c 1 -o gpurun_out/prof_kb_update_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 2 --warm 150 --dict-cap 128 --resident > gpurun_out/kbrl_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 150 -c 1 -o gpurun_out/prof_kb_predict_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 2 --warm 150 --dict-cap 128 --resident >> gpurun_out/kbrl_ncu.log 2>&1
python tools/kbrl_loop.py --envs 16384 --steps 50 --warm 150 --dict-cap 128 --resident > gpurun_out/kbrl_loop_resident_16384_warm150.json 2>> gpurun_out/bench.err; cat gpurun_out/kbrl_loop_resident_16384_warm150.json
ls -la gpurun_out | tail -8
