timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q --tb=line 2>&1 | tail -3
timeout 300 python scripts/prof_fwd.py > /dev/null 2>&1 && NFWD=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_fwd.csv python scripts/prof_fwd.py > gpurun_out/ncu_fwd.log 2>&1
echo ncu exit $?
