mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/tests_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/tests_tc.log
tail -30 gpurun_out/tests_tc.log
if grep -q "pytest tc exit 0" gpurun_out/tests_tc.log; then
  SNB200_CONV=tc3 timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --tb=short -rP > gpurun_out/tests_e2e_tc3.log 2>&1
  echo "pytest e2e tc3 exit $?" >> gpurun_out/tests_e2e_tc3.log
  grep "parity\]" gpurun_out/tests_e2e_tc3.log | tail -40; tail -3 gpurun_out/tests_e2e_tc3.log
  SNB200_CONV=tc3 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tc3.log 2>&1; tail -1 gpurun_out/bench_tc3.log
  SNB200_CONV=tc1 timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_tc1.log 2>&1; tail -1 gpurun_out/bench_tc1.log
fi
