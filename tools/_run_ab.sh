mkdir -p gpurun_out
echo "=== sustained"
timeout 120 python tools/gpu_sustained_gemm.py 1.2 ours 2>&1 | tail -12
timeout 120 python tools/gpu_power_diag.py 3 2>&1 | tail -10
echo "=== tests"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== bench"
timeout 300 python bench.py > gpurun_out/s3r_bench.json 2>gpurun_out/s3r_bench.err; cat gpurun_out/s3r_bench.json | cut -c1-1500
