set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_driver_reference_suite.py tests/test_gpu_driver.py -q --durations=5 ) > gpurun_out/pytest_refsuite.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_refsuite.log; tail -25 gpurun_out/pytest_refsuite.log
