set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "from metsim_b200 import build; print('sources', build.source_hash(), 'built', build.built_hash())"
timeout 600 python profiles/tools/timeline.py --steps 2 --out gpurun_out/timeline_c3.txt > gpurun_out/timeline_c3.log 2>&1; echo "rc=$?"
tail -60 gpurun_out/timeline_c3.log
timeout 300 python profiles/tools/timeline.py --workload global_half_deg --steps 2 --out gpurun_out/timeline_c2.txt > gpurun_out/timeline_c2.log 2>&1; echo "rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['breakdown_ms']); print(json.dumps(d['e2e'])[:700])
PY
