#!/bin/bash
# Round-2 session B: tests, the bench line with all side measurements, reference arm, the 2000-step config-3 protocol,
# compute-sanitizer (memcheck + racecheck), ncu launch list + full captures.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_scn0_65536.json 2> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_scn0_65536.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_reference.json
timeout 900 python tools/kbrl_loop.py --envs 16384 --steps 1980 --warm 20 --report 200,1000,2000 --resident > gpurun_out/kbrl_loop_2000.json 2>> gpurun_out/bench.err; cat gpurun_out/kbrl_loop_2000.json; tail -3 gpurun_out/bench.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py 30 > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.txt; tail -8 gpurun_out/sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py 12 > gpurun_out/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.txt; tail -8 gpurun_out/sanitizer_racecheck.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4210 -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --burn-in 600 --no-cpu-baseline --no-configs > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embb_step_smem -s 600 -c 1 -o gpurun_out/prof_smem_65536 -f python tools/ncu_step.py --envs 65536 --burn-in 600 --steps 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmtc_scan_kernel -s 600 -c 1 -o gpurun_out/prof_mmtc_scan_65536 -f python tools/ncu_step.py --scenario 3 --envs 65536 --burn-in 600 --steps 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 1600 -c 2 -o gpurun_out/prof_kb_update_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 5 --warm 800 --resident > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 1600 -c 2 -o gpurun_out/prof_kb_predict_16384 -f python tools/kbrl_loop.py --envs 16384 --steps 5 --warm 800 --resident > /dev/null 2>&1
ls -la gpurun_out | tail -16
