timeout 600 python -m pytest tests/test_gpu_conv_layers.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for np in 1 0; do echo "== RYK_NO_PDL=$np"; RYK_NO_PDL=$np timeout 600 python bench.py --steps 40 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','rtf','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_step_in_kernel'])"; done
