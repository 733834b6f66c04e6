mkdir -p gpurun_out
timeout 300 python scripts/time_conv2d.py 2>&1 | grep -E "passes3 tmemA"
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_e2e.py tests/test_gpu_backward.py -m gpu -q --tb=short -x 2>&1 | tail -3
timeout 900 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
print("fwd", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocking ms", round(d["e2e"]["blocking_call_ms"],3), "adapt", round(d["adapt"]["value"],2), "steps/s", round(d["adapt"]["ms_per_step"],2), "ms")
PY
