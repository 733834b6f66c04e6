timeout 120 python scripts/time_conv3d.py 2>&1 | tail -22
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_e2e.py tests/test_gpu_backward.py -m gpu -q --tb=short -x 2>&1 | tail -5
