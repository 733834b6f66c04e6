mkdir -p gpurun_out
export SNB200_CONV=tc3
python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/plain_prof.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc3.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/plain_prof2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_c32_tc_kernel -s 12 -c 2 -o gpurun_out/prof_tc3 \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail
