# stream-priority sweep of the session pipeline: RYK_PRIO = levels of (gate, analysis, stage 1, stage 2, synthesis) above the lowest
for p in "5,5,5,0,5" "0,0,0,5,0" "5,0,0,5,0" "5,0,5,5,0" "5,2,3,5,2" "5,5,5,5,5" "5,3,4,5,3"; do
  echo -n "RYK_PRIO=$p: "
  RYK_PRIO=$p RYK_STAGE_TIMES=1 timeout 300 python bench.py --steps 60 2>&1 | tail -1 | python -c "
import sys, json, numpy as np
d = json.loads(sys.stdin.read())
st = d['stage_timeline']; s = np.array(st['start_ms']).reshape(-1, 5); e = np.array(st['end_ms']).reshape(-1, 5)
print(round(d['value']), 'chunks/s, e2e', round(d['e2e']['value']), 'stage ms', np.round((e - s).mean(0), 3).tolist(), 'roofline', round(d['roofline']['achieved']), 'host', round(d['host_enqueue_ms_per_step'], 3))"
done
