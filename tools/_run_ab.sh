mkdir -p gpurun_out
echo "=== tests"
timeout 300 python -m pytest tests/test_gpu_poolscan.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/prof_kernels.py sim 2>&1 | tail -7
