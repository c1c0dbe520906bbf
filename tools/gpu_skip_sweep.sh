# which co-running stage slows stage 2 down?  RYK_SESSION_SKIP bits: 1 analysis, 2 stage 1, 4 stage-2 k4 layers, 8 synthesis (results are garbage; timing only)
for m in 0 1 2 8 9 11 4; do
  echo -n "RYK_SESSION_SKIP=$m: "
  RYK_SESSION_SKIP=$m RYK_HOST_PROF=1 RYK_STAGE_TIMES=1 timeout 300 python bench.py --steps 60 2>gpurun_out/skip_$m.err | tail -1 | python -c "
import sys, json, numpy as np
d = json.loads(sys.stdin.read())
st = d['stage_timeline']; s = np.array(st['start_ms']).reshape(-1, 5); e = np.array(st['end_ms']).reshape(-1, 5)
print(round(d['value']), 'chunks/s, e2e', round(d['e2e']['value']), 'stage ms', np.round((e - s).mean(0), 3).tolist(), 'roofline', d['roofline']['achieved'] and round(d['roofline']['achieved']), 'host', round(d['host_enqueue_ms_per_step'], 3))"
  grep "host prof" gpurun_out/skip_$m.err
done
