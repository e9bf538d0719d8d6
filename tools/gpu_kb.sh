#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kbrl.py -m gpu -x -q --timeout 900 2>&1 | tail -3
for hv in 1000; do
timeout 900 python tools/kbrl_loop.py --envs 16384 --steps 1980 --warm 20 --report 200,1000,2000 --resident --dict-cap 2048 --heavy $hv > gpurun_out/kbrl_loop_2000_h$hv.json 2> gpurun_out/kbrl.err; python -c "
import json; k=json.load(open('gpurun_out/kbrl_loop_2000_h$hv.json')); print('heavy $hv:', k['env_steps_per_s'], k['ms_per_step_wall'])
for c in k['checkpoints']: print({x:c[x] for x in ('after_steps','ms_env','ms_update_control','ms_select_action','dict_mean','dict_max','cap_hits','digest_sizes')})"; tail -2 gpurun_out/kbrl.err
done
