for v in "RYK_PRIO=5,5,5,5,5" "RYK_PRIO=0,0,0,5,0" "RYK_PRIO=5,2,5,4,5" "RYK_PRIO=5,5,5,0,5"; do
  env $v timeout 200 python bench.py --steps 200 --warmup 5 --no-extra --sustain 0 > gpurun_out/r2t_bench.log 2>&1
  python - "$v" <<PY
import sys, json
for ln in open("gpurun_out/r2t_bench.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "TF", round(d["roofline"]["achieved"]), "ms_in_kernel", round(d["roofline"]["ms_per_step_in_kernel"], 4))
    elif "rror" in ln:
        print(ln[:300])
PY
done
