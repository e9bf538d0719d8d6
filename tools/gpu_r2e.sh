#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
bash tools/sweep_warp.sh base
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/kb_crossover tools/kb_crossover.cu && timeout 600 /tmp/kb_crossover > gpurun_out/kb_crossover.jsonl; cat gpurun_out/kb_crossover.jsonl
