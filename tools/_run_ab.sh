timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -x -q -k fused -s 2>&1 | grep "step loss\|assert\|^E" | head -20
