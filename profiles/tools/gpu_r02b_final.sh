# last check of the round: smoke() and the whole GPU suite on the final tree
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log; tail -8 gpurun_out/pytest_final.log
