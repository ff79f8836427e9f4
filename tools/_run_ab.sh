mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s4a.json 2>/dev/null; cut -c1-200 gpurun_out/s4a.json
timeout 120 python tools/gpu_power_diag.py 3 983 2>&1 | tail -10
