mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kbrl.py tests/test_wrappers.py tests/test_gym_shim.py -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu_kbrl.txt; cat gpurun_out/pytest_gpu_kbrl.txt
timeout 900 python tools/kbrl_loop.py --envs 16384 --steps 1980 --warm 20 --report 200,1000,1500,2000 --resident --dict-cap 2048 > gpurun_out/kbrl_loop_2000.json 2>gpurun_out/kbrl.err; cat gpurun_out/kbrl_loop_2000.json; tail -3 gpurun_out/kbrl.err
timeout 600 python bench.py --steps 20 --warmup 5 --scenario 3 --no-cpu-baseline --no-configs > gpurun_out/bench_scn3.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_scn3.json')); print(d['value'], d['ms_per_step'], d['roofline']['side_kernels_ms'], d['roofline'].get('mmtc_scan'))"
