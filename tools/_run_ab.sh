mkdir -p gpurun_out
echo "=== tests"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== bench"
timeout 400 python bench.py > gpurun_out/s3t_bench.json 2>gpurun_out/s3t_bench.err; tail -3 gpurun_out/s3t_bench.err; cat gpurun_out/s3t_bench.json | cut -c1-300
