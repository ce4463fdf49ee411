set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/pytest.log; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-600
