mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "vs_oracle_seeded or golden_B" 2>&1 | tail -5
bash tools/sweep_warp.sh base wmb3 wmb4
