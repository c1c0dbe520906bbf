# Round 2 (second session) final evidence run, one GPU: full GPU suite (soak excluded: it ran green in call 1 and the driver re-runs everything),
# default bench line, fused stage-1 timeline, ncu launch list + full capture of one pipelined step, compute-sanitizer on the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not soak" --durations=8 > gpurun_out/f_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -14 gpurun_out/f_gpu_suite.log
python -m pytest tests/test_gpu_crepe.py tests/test_gpu_harvest.py tests/test_gpu_s1_fused.py -q -s > gpurun_out/f_new_rows_parity.log 2>&1; echo "new rows rc=$?"; grep -E "crepe|end to end|harvest session|stage1 T=260|passed|failed" gpurun_out/f_new_rows_parity.log | head -20
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'], 'launches', d.get('gpu_launches'), 'steps', d['steps'])
print('sustained', {k: d['sustained'][k] for k in ('value', 'seconds', 'stage2_tflops')} if 'sustained' in d else None)
print('extras', {k: (round(v['value'], 1), v.get('stage2_tflops')) for k, v in d.get('extra_configs', {}).items()})
print('cpu_baseline', d.get('cpu_baseline'))
PY
python tools/gpu_s1_bench.py 128 256 384 512 640 > gpurun_out/f_s1_bench.txt 2>&1; echo "s1 bench rc=$?"; grep "Tp" gpurun_out/f_s1_bench.txt
export RYK_HOST_BUCKETS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 4 --warmup 3 --no-extra --sustain 0 > gpurun_out/r02b_launches_bench.log 2>&1
N=$(grep -c "gpu__time_duration" gpurun_out/r02b_launches.csv); PER=$(( N / 7 )); echo "launches total $N per step $PER"
python tools/launch_summary.py gpurun_out/r02b_launches.csv $PER 4 > gpurun_out/r02b_launches_summary.txt 2>&1; head -24 gpurun_out/r02b_launches_summary.txt
SKIP=$(( N - PER ))
timeout 600 ncu --set full --clock-control none --import-source on -s $SKIP -c $PER -f -o /tmp/r02b_full python bench.py --steps 4 --warmup 3 --no-extra --sustain 0 > gpurun_out/r02b_full_bench.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py /tmp/r02b_full.ncu-rep gpurun_out/r02b_ncu_full_one_step.csv > /dev/null 2>&1; wc -l gpurun_out/r02b_ncu_full_one_step.csv
unset RYK_HOST_BUCKETS
export RYK_TEST_QUICK=1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_s1_fused.py "tests/test_gpu_harvest.py::test_harvest_matches_oracle_stage_by_stage[0.3-4]" "tests/test_gpu_crepe.py::test_crepe_network_and_decoders_match_oracle[tiny-0.9--2.0]" -x -q > gpurun_out/r02b_compute_sanitizer_memcheck_new_kernels.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02b_compute_sanitizer_memcheck_new_kernels.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_s1_fused.py "tests/test_gpu_harvest.py::test_harvest_matches_oracle_stage_by_stage[0.3-4]" -x -q > gpurun_out/r02b_compute_sanitizer_racecheck_new_kernels.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02b_compute_sanitizer_racecheck_new_kernels.log
