mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi1.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -rP > gpurun_out/tests_kernels.log 2>&1
echo "pytest kernels exit $?" >> gpurun_out/tests_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --tb=short -rP > gpurun_out/tests_e2e.log 2>&1
echo "pytest e2e exit $?" >> gpurun_out/tests_e2e.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench1.log
tail -4 gpurun_out/tests_kernels.log; tail -4 gpurun_out/tests_e2e.log; tail -3 gpurun_out/bench1.log
