mkdir -p gpurun_out
python -m pytest tests/test_gpu_crepe.py tests/test_gpu_harvest.py -q -s > gpurun_out/c6_rows.log 2>&1; echo "rows rc=$?"; grep -E "crepe|voicing states|end to end|harvest session|decimated|passed|failed" gpurun_out/c6_rows.log | head -30
python bench.py > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c6_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'launches', d.get('gpu_launches'), 'steps', d['steps'])
print('sustained', {k: d['sustained'][k] for k in ('value', 'seconds', 'stage2_tflops')} if 'sustained' in d else None)
print('extras', {k: (round(v['value'], 1), v.get('stage2_tflops')) for k, v in d.get('extra_configs', {}).items()})
PY
python bench.py --steps 100 --warmup 5 --no-extra --sustain 0 > gpurun_out/c6_bench_k100.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/c6_bench_k100.json').read().strip().splitlines()[-1]); print('K=100 value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])"
