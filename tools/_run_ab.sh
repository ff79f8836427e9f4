mkdir -p gpurun_out
echo "=== towers tests (u8 LUT)"
timeout 200 python -m pytest tests/test_gpu_towers.py -m gpu -x -q 2>&1 | tail -2
echo "=== bench N=2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r01f_bench_n2.json 2> gpurun_out/r01f_bench_n2.err; tail -2 gpurun_out/r01f_bench_n2.err; cut -c1-400 gpurun_out/r01f_bench_n2.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
