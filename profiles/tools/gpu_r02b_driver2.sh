# the driver's GPU tests (with the cell ranges over several contexts) and files-to-files runs: one call, two ranges on one device
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_driver.py -q > gpurun_out/pytest_driver.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_driver.log; tail -5 gpurun_out/pytest_driver.log
timeout 200 python profiles/tools/driver_scale.py > gpurun_out/driver_scale_1.log 2> gpurun_out/driver_scale_1.err; cat gpurun_out/driver_scale_1.log; tail -2 gpurun_out/driver_scale_1.err
timeout 200 python profiles/tools/driver_scale.py --devices 0,0 > gpurun_out/driver_scale_00.log 2> gpurun_out/driver_scale_00.err; cat gpurun_out/driver_scale_00.log; tail -2 gpurun_out/driver_scale_00.err
