mkdir -p gpurun_out
python scripts/prof_adapt.py > gpurun_out/prof_adapt_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_adapt.csv \
    python scripts/prof_adapt.py > gpurun_out/ncu_adapt.log 2>&1
echo "ncu exit $?"
SNB200_CONV=tc3 timeout 600 python bench.py --steps 30 --warmup 5 --skip-adapt --skip-cpu > gpurun_out/bench_fwd.log 2>&1; tail -c 600 gpurun_out/bench_fwd.log
