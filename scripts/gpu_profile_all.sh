mkdir -p gpurun_out
timeout 300 python scripts/prof_fwd.py > gpurun_out/prof_fwd_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_fwd.csv python scripts/prof_fwd.py > gpurun_out/ncu_fwd.log 2>&1
echo "ncu fwd exit $?"
NSTEPS=3 timeout 300 python scripts/prof_adapt.py > gpurun_out/prof_adapt_plain.log 2>&1 && \
NSTEPS=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_adapt.csv python scripts/prof_adapt.py > gpurun_out/ncu_adapt.log 2>&1
echo "ncu adapt exit $?"
timeout 300 python scripts/prof_top.py > gpurun_out/prof_top_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv2d_c32_tc_kernel|conv3d_c32_tma_kernel|conv_c32_wgrad_tc_kernel|cost_volume_fwd_kernel|conv_c32_taps_tc_kernel|tapsum_softargmin_kernel|conv_small_tc_kernel" -s 22 -c 11 -o gpurun_out/prof_top python scripts/prof_top.py > gpurun_out/ncu_top.log 2>&1
echo "ncu top exit $?"
