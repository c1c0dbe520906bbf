mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c10_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c10_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c10_bench_2gpu.json 2> gpurun_out/c10_bench_2gpu.err; echo "2gpu rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c10_bench_2gpu.json').read().strip().splitlines()[-1])
print('2 GPUs: value', d['value'], 'e2e', d['e2e']['value'], 'n_gpus', d['n_gpus'], 'extras', {k: round(v['value'], 1) for k, v in d.get('extra_configs', {}).items()})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 3 > gpurun_out/c10_ref_2gpu.json 2> gpurun_out/c10_ref_2gpu.err; echo "ref rc=$?"; tail -1 gpurun_out/c10_ref_2gpu.json | cut -c1-300
