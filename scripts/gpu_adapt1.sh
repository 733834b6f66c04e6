mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
print("fwd", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "adapt", round(d["adapt"]["value"],2), "steps/s", round(d["adapt"]["ms_per_step"],2), "ms")
for k,v in d["kernels"].items(): print(k, round(v["ms"]*1000,1),"us", round(v["achieved"],1), v["unit"], round(v["frac"],3))
PY
NSTEPS=2 timeout 300 python scripts/prof_adapt.py > gpurun_out/prof_adapt_plain.log 2>&1 && \
NSTEPS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_adapt.csv python scripts/prof_adapt.py > gpurun_out/ncu_adapt.log 2>&1
echo "ncu exit $?"
