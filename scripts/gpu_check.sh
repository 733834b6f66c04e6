# One GPU call: the whole `-m gpu` suite file by file, then a short bench.  Usage: gpurun -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
for f in test_gpu_kernels test_gpu_conv_tc test_gpu_wgrad_tc test_gpu_e2e test_gpu_backward; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --tb=short > gpurun_out/$f.log 2>&1
  echo "$f exit $?: $(grep -E 'passed|failed' gpurun_out/$f.log | tail -1)"
done
timeout 900 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
print("fwd", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocking ms", round(d["e2e"]["blocking_call_ms"],3), "adapt", round(d["adapt"]["value"],2), "steps/s", round(d["adapt"]["ms_per_step"],2), "ms")
for k,v in d["kernels"].items(): print(k, round(v["ms"]*1000,1),"us", round(v["achieved"],1), v["unit"], round(v["frac"],3))
PY
