mkdir -p gpurun_out
python scripts/prof_tc.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:c32_tc_kernel -c 6 -o gpurun_out/prof_tc_big \
    python scripts/prof_tc.py > gpurun_out/ncu_full2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full2.log
