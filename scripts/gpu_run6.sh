mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/test_gpu_conv_tc.log 2>&1
echo "tc exit $?"; grep -E "passed|failed" gpurun_out/test_gpu_conv_tc.log | tail -1; grep -E "Error|assert" gpurun_out/test_gpu_conv_tc.log | head -5
if grep -q " passed" gpurun_out/test_gpu_conv_tc.log && ! grep -q "failed" gpurun_out/test_gpu_conv_tc.log; then
for f in test_gpu_e2e test_gpu_backward; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q --tb=short > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1
done
timeout 900 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
print("fwd", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "adapt", round(d["adapt"]["value"],2), "steps/s", round(d["adapt"]["ms_per_step"],2), "ms")
for k,v in d["kernels"].items(): print(k, round(v["ms"]*1000,1),"us", round(v["achieved"],1), v["unit"], round(v["frac"],3))
PY
fi
