timeout 600 python -m pytest tests/test_gpu_conv_layers.py -x -q 2>&1 | tail -3
timeout 300 python tools/gpu_bench_layers.py 2>&1 | tail -18
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 600 python bench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','rtf','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_step_in_kernel'])"
