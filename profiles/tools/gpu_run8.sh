set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x -k "fast_path or fuzz or benchmarked or golden or precipitation or subsets or unit_conv or solar" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python profiles/tools/probe_solar.py > gpurun_out/probe_solar.log 2>&1
K='regex:thermo|streams_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 16 --csv --log-file gpurun_out/launches_pair.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list.log 2>&1
MSB_STREAMS_SPLIT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 24 --csv --log-file gpurun_out/launches_split.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
MSB_STREAMS_SPLIT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-verify > gpurun_out/bench_split.log 2> gpurun_out/bench_split.err
tail -3 gpurun_out/pytest.log; cat gpurun_out/probe_solar.log; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-300; cat gpurun_out/bench_split.log | cut -c1-300
