mkdir -p gpurun_out
timeout 300 python scripts/time_conv2d.py 2>&1 | head -30
timeout 200 python scripts/tc2d_counters.py 2>&1 | grep "dil1" | head -8
