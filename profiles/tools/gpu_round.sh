set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "from metsim_b200 import build; print('sources', build.source_hash(), 'built', build.built_hash())"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 400 python profiles/tools/timeline.py --steps 2 > gpurun_out/timeline_c3.log 2>&1; echo "rc=$?"
timeout 300 python profiles/tools/timeline.py --workload global_half_deg --steps 2 > gpurun_out/timeline_c2.log 2>&1; echo "rc=$?"
MSB_SW_TABLE=1 timeout 300 python profiles/tools/timeline.py --workload global_half_deg --steps 2 > gpurun_out/timeline_c2_plain.log 2>&1; echo "rc=$?"
MSB_SW_TABLE=2 timeout 400 python profiles/tools/timeline.py --steps 2 > gpurun_out/timeline_c3_win.log 2>&1; echo "rc=$?"
timeout 400 python bench.py --workload global_half_deg --steps 10 --warmup 3 > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
grep -h "events\|ms/step" gpurun_out/timeline_c3.log gpurun_out/timeline_c2.log gpurun_out/timeline_c2_plain.log gpurun_out/timeline_c3_win.log | grep -v "Memset\|reduce\|daily_kernel"
cut -c1-400 gpurun_out/bench_c2.log
