#!/bin/bash
# TMA-staged fading-trace columns in the warp-per-unit kernel (build: RS_BUILD_TAG=tma RS_NVCC_EXTRA=-DRS_WARP_TMA): parity, then timing vs the L2-load build
mkdir -p gpurun_out
export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_tma.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "(vs_oracle_seeded or golden_B) and 3" 2>&1 | tail -4
unset RS_B200_LIB
bash tools/sweep_warp.sh base tma
