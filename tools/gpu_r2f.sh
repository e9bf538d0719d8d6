#!/bin/bash
# Round-2 final evidence set: tests, smoke, bench line (with configs[]), reference arm, config-3 protocol, multiplexed mode,
# launch list + full captures of the side kernels.  Outputs -> gpurun_out/ (summaries copied to profiles/r02f_*)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_scn0_65536.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_scn0_65536.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_reference.json
timeout 900 python tools/kbrl_loop.py --envs 16384 --steps 1980 --warm 20 --report 200,1000,2000 --resident --dict-cap 2048 > gpurun_out/kbrl_loop_2000.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/kbrl_loop_2000.json
timeout 300 python tools/mux_bench.py > gpurun_out/mux_bench.txt 2>&1; tail -1 gpurun_out/mux_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4210 -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --burn-in 600 --no-cpu-baseline --no-configs > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmtc_scan_kernel -s 600 -c 1 -o gpurun_out/prof_mmtc_scan_65536 -f python tools/ncu_step.py --scenario 3 --envs 65536 --burn-in 600 --steps 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update -s 3000 -c 2 -o gpurun_out/prof_kb_update_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 5 --warm 1500 --resident --dict-cap 2048 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict -s 4500 -c 3 -o gpurun_out/prof_kb_predict_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 5 --warm 1500 --resident --dict-cap 2048 > /dev/null 2>&1
ls -la gpurun_out | tail -12
