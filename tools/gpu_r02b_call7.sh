mkdir -p gpurun_out
RYK_STAGE_TIMES=1 python bench.py --streams-per-gpu 8 --buffer-time 1.0 --steps 12 --warmup 4 --no-extra --sustain 0 > gpurun_out/c7_group_stage_times.json 2> gpurun_out/c7_group.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c7_group_stage_times.json').read().strip().splitlines()[-1])
print('group 8x1.0s value', d['value'], 'e2e', d['e2e']['value'], 'ms_per_step', d['ms_per_step'], 'roofline', d['roofline']['achieved'], 'host_enq', d.get('host_enqueue_ms_per_step'))
tl = d.get('stage_timeline')
if tl:
    for s, e in zip(tl['start_ms'], tl['end_ms']):
        print(' '.join(f'{n}:{a:.2f}-{b:.2f}' for n, a, b in zip(['gate', 'ana', 's1', 's2', 'syn'], s, e)))
PY
RYK_STAGE_TIMES=1 python bench.py --buffer-time 1.0 --steps 12 --warmup 4 --no-extra --sustain 0 > gpurun_out/c7_single_1s_stage_times.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c7_single_1s_stage_times.json').read().strip().splitlines()[-1])
print('single 1.0s value', d['value'], 'ms_per_step', d['ms_per_step'], 'roofline', d['roofline']['achieved'])
tl = d.get('stage_timeline')
if tl:
    for s, e in zip(tl['start_ms'][-4:], tl['end_ms'][-4:]):
        print(' '.join(f'{n}:{a:.2f}-{b:.2f}' for n, a, b in zip(['gate', 'ana', 's1', 's2', 'syn'], s, e)))
PY
