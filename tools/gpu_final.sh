#!/bin/bash
# final validation of a round: full GPU suite, smoke, the bench line (with configs[], CPU arms), the reference arm, the launch
# list of the bench command, the 2000-step KBRL protocol and the multiplexed-L1 number.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_scn0_65536.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_scn0_65536.json')); print(d['value'], d['e2e']['value'], d['e2e']['blocking_value'], d['roofline']['frac'], d.get('issue_roofline',{}).get('frac'))
for c in d['configs']: print(c['config'][:40], c['value'], c.get('ms_per_step'), c.get('ms_env'))"; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_scn0_65536.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 python tools/kbrl_loop.py --envs 16384 --steps 1980 --warm 20 --report 200,1000,1500,2000 --resident --dict-cap 2048 > gpurun_out/kbrl_loop_2000_cap2048.json 2>> gpurun_out/bench.err; python -c "
import json; k=json.load(open('gpurun_out/kbrl_loop_2000_cap2048.json'))
for c in k['checkpoints']: print('step %d: env %.3f update %.3f select %.3f  D mean %.1f p99 %.0f max %d cap_hits %d pool_hits %d digest %s' % (c['after_steps'], c['ms_env'], c['ms_update_control'], c['ms_select_action'], c['dict_mean'], c['dict_p99'], c['dict_max'], c['cap_hits'], c['pool_hits'], c['digest_sizes']))
print('overall %.4g env-steps/s' % k['env_steps_per_s'])"
timeout 300 python tools/mux_bench.py 2>> gpurun_out/bench.err | tee gpurun_out/mux_bench.txt
