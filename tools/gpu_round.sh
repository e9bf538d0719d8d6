#!/bin/bash
# One GPU evidence session: tests, smoke, bench lines of every BASELINE config shape, ncu launch lists + full captures of
# the dominant kernels.  Outputs -> gpurun_out/ (summarise with tools/ncu_*.py and copy what is quoted into profiles/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_scn0_65536.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_scn0_65536.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --envs-per-gpu 4096 --no-cpu-baseline > gpurun_out/bench_scn0_4096.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_scn0_4096.json
python bench.py --steps 20 --warmup 5 --scenario 3 --no-cpu-baseline > gpurun_out/bench_scn3_65536.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_scn3_65536.json
python tools/kbrl_loop.py --envs 16384 --steps 30 --warm 170 --dict-cap 128 --resident > gpurun_out/kbrl_loop_resident_16384.json 2>> gpurun_out/bench.err; cat gpurun_out/kbrl_loop_resident_16384.json
python tools/kbrl_loop.py --envs 16384 --steps 5 --warm 30 --dict-cap 128 > gpurun_out/kbrl_loop_host_16384.json 2>> gpurun_out/bench.err; cat gpurun_out/kbrl_loop_host_16384.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 2420 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --burn-in 600 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:embb_step_smem -s 600 -c 1 -o gpurun_out/prof_smem_65536 -f python tools/ncu_step.py --envs 65536 --burn-in 600 --steps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mmtc_scan_kernel -s 600 -c 1 -o gpurun_out/prof_mmtc_scan_65536 -f python tools/ncu_step.py --scenario 3 --envs 65536 --burn-in 600 --steps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 170 -c 1 -o gpurun_out/prof_kb_update_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 2 --warm 170 --dict-cap 128 --resident > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 170 -c 1 -o gpurun_out/prof_kb_predict_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 2 --warm 170 --dict-cap 128 --resident > /dev/null 2>&1
ls -la gpurun_out | tail -14
