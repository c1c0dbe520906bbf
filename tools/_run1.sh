timeout 600 python -m pytest tests/test_gpu_conv_layers.py -x -q 2>&1 | tail -5
timeout 300 python tools/gpu_bench_layers.py 2>&1 | tail -4
