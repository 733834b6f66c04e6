mkdir -p gpurun_out
for f in test_gpu_kernels test_gpu_conv_tc test_gpu_wgrad_tc test_gpu_e2e test_gpu_backward; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --tb=short > gpurun_out/$f.log 2>&1
  echo "$f exit $?: $(grep -E 'passed|failed' gpurun_out/$f.log | tail -1)"
done
