# Round-2 evidence (run under gpurun, one GPU): launch list + `--set full` of one pipelined step (host-bucket path: ncu does not see
# kernels inside CUDA conditional-graph bodies, so the stage-1 switch is replaced by its host-selected equivalent for profiling only).
export RYK_HOST_BUCKETS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-extra --sustain 0 > gpurun_out/r02_launches_bench.log 2>&1
N=$(grep -c "gpu__time_duration" gpurun_out/r02_launches.csv)
PER=$(( N / 7 ))
echo "launches total $N per step $PER"
python tools/launch_summary.py gpurun_out/r02_launches.csv $PER 4 > gpurun_out/r02_launches_summary.txt 2>&1
SKIP=$(( N - PER ))
ncu --set full --clock-control none --import-source on -s $SKIP -c $PER -f -o /tmp/r02_full python bench.py --steps 4 --warmup 3 --no-extra --sustain 0 > gpurun_out/r02_full_bench.log 2>&1
python tools/ncu_summary.py /tmp/r02_full.ncu-rep gpurun_out/r02_ncu_full_one_step.csv
# source-level stall profile of k_d4c (top lines by samples)
ncu -i /tmp/r02_full.ncu-rep --page source --csv -k regex:k_d4c 2>/dev/null > /tmp/d4c_src.csv
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/d4c_src.csv')))
hdr = None
for i, r in enumerate(rows):
    if 'Source' in r and any('Sampl' in c for c in r):
        hdr = i; break
if hdr is None:
    print('no source table'); raise SystemExit
h = rows[hdr]
si = h.index('Source'); ci = [i for i, c in enumerate(h) if c.startswith('# Samples') or c == 'Samples' or 'Warp Stall Sampling (All' in c]
ci = ci[0] if ci else None
out = []
for r in rows[hdr + 1:]:
    try:
        out.append((float(r[ci].replace(',', '') or 0), r[si][:150], r[0][:40]))
    except Exception:
        pass
out.sort(reverse=True)
tot = sum(o[0] for o in out) or 1
with open('gpurun_out/r02_d4c_source_hot_lines.txt', 'w') as f:
    f.write('share  samples  line  source\n')
    for s, src, ln in out[:40]:
        f.write(f'{s / tot:6.1%} {s:8.0f}  {ln}  {src}\n')
print(open('gpurun_out/r02_d4c_source_hot_lines.txt').read()[:3000])
PY
