set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
for R in 1 2 4 8; do
  timeout 300 python bench.py --workload global_half_deg --patch-rows $R --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c2_rows$R.log 2> gpurun_out/bench_c2_rows$R.err
done
K='regex:solar|daily|thermo|streams|sw_day|prec_day|disagg|lw_rescale'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:sw_day|prec_day|solar_rows' -s 2 -c 3 -o gpurun_out/r02c_streams python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest.log; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-300
