set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x -k "solar or golden or fast_path or fuzz or benchmarked" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
K='regex:solar|sw_day'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 12 --csv --log-file gpurun_out/launches_pf1.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list.log 2>&1
MSB_DAY_PF=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 12 --csv --log-file gpurun_out/launches_pf2.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
MSB_DAY_PF=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_pf2.log 2> gpurun_out/bench_pf2.err
tail -3 gpurun_out/pytest.log; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-300; cat gpurun_out/bench_pf2.log | cut -c1-300
