timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
RYK_BENCH_TRACE=1 RYK_STAGE_TIMES=1 timeout 600 python bench.py --steps 40 2>&1 | grep -E "host us|metric" | python -c "
import sys,json,numpy as np
for l in sys.stdin:
    if l.startswith('host'): print(l.strip())
    else:
        d=json.loads(l); print(round(d['ms_per_step']*1000,1), 'us/step', d['value'], d['e2e']['value'])
        t=d['stage_timeline']; s=np.array(t['start_ms'])*1000; e=np.array(t['end_ms'])*1000
        for i in range(len(s)):
            print('step',i,' '.join(f'{n[:6]}:[{a:7.0f},{b:7.0f}]' for n,a,b in zip(t['stages'],s[i],e[i])))
"
timeout 600 python bench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','rtf','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])"
