mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --tb=line -k golden 2>&1 | tail -3
NSTEPS=2 timeout 300 python scripts/prof_adapt.py > gpurun_out/prof_adapt_plain.log 2>&1 && \
NSTEPS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_adapt.csv python scripts/prof_adapt.py > gpurun_out/ncu_adapt.log 2>&1
echo "ncu adapt exit $?"
