timeout 300 python scripts/time_conv2d.py 2>&1 | grep -E "passes3"
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_e2e.py tests/test_gpu_backward.py -m gpu -q --tb=line -x 2>&1 | tail -4
timeout 200 python scripts/tc2d_counters.py 2>&1 | grep "dil1 passes=3 tmemA"
