mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wgrad_tc.py -m gpu -q --tb=short > gpurun_out/test_gpu_wgrad_tc.log 2>&1
echo "wgrad exit $?"; tail -5 gpurun_out/test_gpu_wgrad_tc.log
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --tb=short > gpurun_out/test_gpu_backward.log 2>&1
echo "backward exit $?"; tail -5 gpurun_out/test_gpu_backward.log
timeout 300 python scripts/time_wgrad.py > gpurun_out/time_wgrad.log 2>&1; cat gpurun_out/time_wgrad.log
