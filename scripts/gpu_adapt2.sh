mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --tb=short > gpurun_out/test_gpu_backward.log 2>&1
echo "backward exit $?"; tail -15 gpurun_out/test_gpu_backward.log
timeout 900 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/bench.log 2>&1; tail -c 1500 gpurun_out/bench.log | head -c 400; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
print("fwd", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "adapt", round(d["adapt"]["value"],2), "steps/s", round(d["adapt"]["ms_per_step"],2), "ms")
PY
