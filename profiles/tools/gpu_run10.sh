set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x -k "golden or fast_path or fuzz or benchmarked or subsets" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
K='regex:sw_day'
for PF in 0 1 -1; do
MSB_DAY_PF=$PF timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4 --csv --log-file gpurun_out/launches_pf$PF.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-300
