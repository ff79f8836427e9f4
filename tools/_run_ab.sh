mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/gpu_coop_step.py 16 10 2>&1 | tail -1
timeout 100 python tools/gpu_coop_step.py 1024 100 2>&1 | tail -1
