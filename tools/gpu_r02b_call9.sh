mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not soak" > gpurun_out/c9_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/c9_gpu_suite.log
python -m pytest tests/test_gpu_headline_parity.py -q -s -k "group" > gpurun_out/c9_group_parity.log 2>&1; grep -E "HEADLINE|group|rmse|passed|failed" gpurun_out/c9_group_parity.log | head
python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; echo "bench rc=$?"
RYK_BENCH_GROUP_FUSED=1 python bench.py --streams-per-gpu 8 --buffer-time 1.0 --steps 12 --warmup 4 --no-extra --sustain 0 > gpurun_out/c9_group8_fused.json 2>/dev/null
python bench.py --streams-per-gpu 8 --buffer-time 1.0 --steps 12 --warmup 4 --no-extra --sustain 0 > gpurun_out/c9_group8_layered.json 2>/dev/null
python bench.py --streams-per-gpu 8 --buffer-time 0.3 --steps 20 --warmup 4 --no-extra --sustain 0 > gpurun_out/c9_group8x03.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c9_bench.json').read().strip().splitlines()[-1])
print('default: value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'sustained', d.get('sustained', {}).get('value'))
print('extras', {k: (round(v['value'], 1), v.get('stage2_tflops') and round(v['stage2_tflops'])) for k, v in d.get('extra_configs', {}).items()})
for n in ('group8_fused', 'group8_layered', 'group8x03'):
    d = json.loads(open(f'gpurun_out/c9_{n}.json').read().strip().splitlines()[-1])
    print(n, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'stage-2', round(d['roofline']['achieved']), 'TF/s')
PY
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest "tests/test_gpu_parity.py::test_group_batched_stage2_matches_oracle_streams" -x -q > gpurun_out/r02b_compute_sanitizer_racecheck_group_32_connections.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02b_compute_sanitizer_racecheck_group_32_connections.log
