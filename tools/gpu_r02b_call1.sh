# Round 2 (second session), GPU call 1: the fused stage-1 kernel -- parity (vs layered path and oracle), the session / headline suites
# that now run through it, and an A/B of the pipelined bench.  Everything goes to gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_s1_fused.py -x -q -s > gpurun_out/c1_s1_fused.log 2>&1; echo "s1_fused rc=$?"
tail -14 gpurun_out/c1_s1_fused.log
timeout 900 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/c1_gpu_suite.log 2>&1; echo "suite rc=$?"
tail -22 gpurun_out/c1_gpu_suite.log
python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c1_bench_fused.json 2> gpurun_out/c1_bench_fused.err; echo "bench fused rc=$?"
RYK_S1_FUSED=0 python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c1_bench_layered.json 2> gpurun_out/c1_bench_layered.err; echo "bench layered rc=$?"
RYK_S1_CLUSTER=8 python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c1_bench_fused_c8.json 2> gpurun_out/c1_bench_fused_c8.err; echo "bench fused c8 rc=$?"
python - <<'PY'
import json
for n in ('fused', 'layered', 'fused_c8'):
    try:
        d = json.loads(open(f'gpurun_out/c1_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, d['value'], d['unit'], 'e2e', d['e2e']['value'], 'roofline', d['roofline'].get('frac'), 'launches', d.get('gpu_launches'), 'host_enq', d.get('host_enqueue_ms_per_step'))
    except Exception as ex:
        print(n, 'unreadable', ex)
PY
RYK_STAGE_TIMES=1 python bench.py --steps 20 --warmup 5 --no-extra --sustain 0 > gpurun_out/c1_bench_stage_times.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/c1_bench_stage_times.json').read().strip().splitlines()[-1]); print('stage_timeline', d.get('stage_timeline'))"
