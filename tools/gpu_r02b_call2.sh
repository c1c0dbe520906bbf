# Round 2 (second session), GPU call 2: restructured fused stage-1 kernel (parity + stand-alone timing + phase timeline + ncu),
# Harvest f0 stage-by-stage parity, pipelined bench A/B.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_s1_fused.py -x -q -s > gpurun_out/c2_s1_fused.log 2>&1; echo "s1_fused rc=$?"; tail -12 gpurun_out/c2_s1_fused.log
python tools/gpu_s1_bench.py > gpurun_out/c2_s1_bench.txt 2>&1; echo "s1 bench rc=$?"; cat gpurun_out/c2_s1_bench.txt
python -m pytest tests/test_gpu_harvest.py -q -s > gpurun_out/c2_harvest.log 2>&1; echo "harvest rc=$?"; grep -v "^$" gpurun_out/c2_harvest.log | tail -60
python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c2_bench_fused.json 2> gpurun_out/c2_bench_fused.err; echo "bench fused rc=$?"
RYK_S1_FUSED=0 python bench.py --steps 40 --warmup 5 --no-extra --sustain 0 > gpurun_out/c2_bench_layered.json 2> gpurun_out/c2_bench_layered.err; echo "bench layered rc=$?"
python - <<'PY'
import json
for n in ('fused', 'layered'):
    try:
        d = json.loads(open(f'gpurun_out/c2_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, d['value'], d['unit'], 'e2e', d['e2e']['value'], 'roofline', d['roofline'].get('frac'), 'launches', d.get('gpu_launches'))
    except Exception as ex:
        print(n, 'unreadable', ex)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s1_fused -s 5 -c 1 -f -o /tmp/c2_s1 python tools/gpu_s1_bench.py 384 > gpurun_out/c2_ncu_s1.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/c2_s1.ncu-rep gpurun_out/c2_ncu_s1_fused.csv > /dev/null 2>&1
ncu -i /tmp/c2_s1.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|Stall|L2 Hit|Throughput|Registers|Issue Slots|Eligible|No Eligible|Warp Cycles Per Issued" | cut -c1-220 | head -40 > gpurun_out/c2_ncu_s1_details.txt; cat gpurun_out/c2_ncu_s1_details.txt | head -30
cp /tmp/c2_s1.ncu-rep gpurun_out/c2_s1.ncu-rep 2>/dev/null; ls -la gpurun_out/c2_s1.ncu-rep
