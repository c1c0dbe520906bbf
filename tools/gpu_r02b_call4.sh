mkdir -p gpurun_out
python -m pytest tests/test_gpu_crepe.py -q -s > gpurun_out/c4_crepe.log 2>&1; echo "crepe rc=$?"; grep -v "^$" gpurun_out/c4_crepe.log | tail -40
python -m pytest tests/test_gpu_s1_fused.py -x -q -s > gpurun_out/c4_s1_fused.log 2>&1; echo "s1_fused rc=$?"; tail -4 gpurun_out/c4_s1_fused.log
python tools/gpu_s1_bench.py 128 384 640 > gpurun_out/c4_s1_bench.txt 2>&1; echo "s1 bench rc=$?"; cat gpurun_out/c4_s1_bench.txt
python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c4_bench_fused.json 2> gpurun_out/c4_bench_fused.err; echo "bench fused rc=$?"
RYK_S1_FUSED=0 python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c4_bench_layered.json 2> gpurun_out/c4_bench_layered.err; echo "bench layered rc=$?"
python - <<'PY'
import json
for n in ('fused', 'layered'):
    try:
        d = json.loads(open(f'gpurun_out/c4_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, d['value'], d['unit'], 'e2e', d['e2e']['value'], 'roofline', d['roofline'].get('frac'), 'launches', d.get('gpu_launches'))
    except Exception as ex:
        print(n, 'unreadable', ex)
PY
