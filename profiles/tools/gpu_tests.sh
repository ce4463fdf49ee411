set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "from metsim_b200 import build; print('sources', build.source_hash(), 'built', build.built_hash())"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log
