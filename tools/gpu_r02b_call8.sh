mkdir -p gpurun_out
for C in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$C RYK_STAGE_TIMES=1 python bench.py --streams-per-gpu 8 --buffer-time 1.0 --steps 12 --warmup 4 --no-extra --sustain 0 > gpurun_out/c8_group_conn$C.json 2>/dev/null
  CUDA_DEVICE_MAX_CONNECTIONS=$C python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c8_single_conn$C.json 2>/dev/null
  CUDA_DEVICE_MAX_CONNECTIONS=$C python bench.py --streams-per-gpu 4 --buffer-time 0.3 --steps 20 --warmup 4 --no-extra --sustain 0 > gpurun_out/c8_group4_conn$C.json 2>/dev/null
done
python - <<'PY'
import json
for C in (8, 32):
    for n in ('group', 'single', 'group4'):
        try:
            d = json.loads(open(f'gpurun_out/c8_{n}_conn{C}.json').read().strip().splitlines()[-1])
            print(f'connections {C:2d} {n:7s}: value {d["value"]:.1f} e2e {d["e2e"]["value"]:.1f} ms/step {d["ms_per_step"]:.3f} stage-2 {d["roofline"]["achieved"]:.0f} TF/s frac {d["roofline"]["frac"]:.3f}')
            tl = d.get('stage_timeline')
            if tl and n == 'group':
                for s, e in list(zip(tl['start_ms'], tl['end_ms']))[-3:]:
                    print('      ' + ' '.join(f'{nm}:{a:.2f}-{b:.2f}' for nm, a, b in zip(['gate', 'ana', 's1', 's2', 'syn'], s, e)))
        except Exception as ex:
            print(C, n, 'unreadable', ex)
PY
