# the evidence run of a round: smoke, tests, ncu --set full (-> traffic JSON of THESE sources), bench (both arms),
# ncu launch list of the SAME bench command, activity timeline, the masked global grid
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
H=$(python -c "from metsim_b200 import build; print(build.source_hash())")
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:solar_rows|solar_moments|daily_fused_kernel|thermo_knot|sw_day|prec_day' -c 6 -o gpurun_out/r02_step_kernels python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_full.log 2>&1
# the first captured call covers the interior days 2..96 of a 97-day tile: 95 days x 24 steps
python profiles/tools/make_traffic.py gpurun_out/r02_step_kernels.ncu-rep 482560 $((95*24)) "$H" r02_step_kernels > /dev/null && cp profiles/disagg_traffic.json gpurun_out/disagg_traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
K='regex:solar|daily|thermo|sw_day|prec_day|disagg|lw_rescale'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_list.log 2>&1
timeout 400 python profiles/tools/timeline.py --steps 2 > gpurun_out/timeline_c3.log 2>&1
timeout 300 python profiles/tools/timeline.py --workload global_half_deg --steps 2 > gpurun_out/timeline_c2.log 2>&1
timeout 400 python bench.py --workload global_half_deg --steps 20 --warmup 5 > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err
tail -3 gpurun_out/pytest.log; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-300
