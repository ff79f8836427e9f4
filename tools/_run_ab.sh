mkdir -p gpurun_out
echo "=== tests"
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_towers.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/gpu_sustained_gemm.py 1.0 ours 2>&1 | tail -12
timeout 120 python tools/gpu_power_diag.py 3 2>&1 | tail -10
