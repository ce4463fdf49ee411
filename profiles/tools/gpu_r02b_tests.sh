# second evidence run of round 2: full GPU suite (with the driver tests) and the driver at a realistic size
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "from metsim_b200 import build; print('sources', build.source_hash(), 'built', build.built_hash())" > gpurun_out/hash.log 2>&1
( time timeout 700 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python profiles/tools/driver_scale.py > gpurun_out/driver_scale.log 2> gpurun_out/driver_scale.err; echo "scale rc=$?" >> gpurun_out/driver_scale.err
tail -15 gpurun_out/pytest.log; cat gpurun_out/driver_scale.log; tail -3 gpurun_out/driver_scale.err
