mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q --tb=short -rP > gpurun_out/test_gpu_backward.log 2>&1
echo "backward exit $?" | tee -a gpurun_out/test_gpu_backward.log
grep -E "passed|failed" gpurun_out/test_gpu_backward.log | tail -1
grep -E "^(FAILED|ERROR)|^____|parity\]|Error|assert " gpurun_out/test_gpu_backward.log | head -60
