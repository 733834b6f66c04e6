mkdir -p gpurun_out
for f in test_gpu_kernels test_gpu_conv_tc test_gpu_e2e; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q --tb=short -rP > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/$f.log
  grep -E "passed|failed" gpurun_out/$f.log | tail -1
done
grep "parity\] kitti" gpurun_out/test_gpu_e2e.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
