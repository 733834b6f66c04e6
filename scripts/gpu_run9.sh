mkdir -p gpurun_out
timeout 300 python scripts/prof_fwd.py > gpurun_out/prof_fwd_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_fwd.csv python scripts/prof_fwd.py > gpurun_out/ncu_fwd.log 2>&1
echo "ncu fwd exit $?"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x 2>&1 | tail -3
