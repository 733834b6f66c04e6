mkdir -p gpurun_out
timeout 300 python scripts/dbg_wgrad.py > gpurun_out/dbg_wgrad.log 2>&1; echo "dbg exit $?"; cat gpurun_out/dbg_wgrad.log | tail -30
timeout 600 python -m pytest tests/test_gpu_wgrad_tc.py -m gpu -q --tb=short > gpurun_out/test_gpu_wgrad_tc.log 2>&1
echo "wgrad exit $?"; tail -15 gpurun_out/test_gpu_wgrad_tc.log
