# files-to-files through the driver on a 2-GPU box: one device, two ranges on two devices, two ranges on one device
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
timeout 200 python profiles/tools/driver_scale.py > gpurun_out/driver_scale_n2_single.log 2> gpurun_out/driver_scale_n2_single.err; cat gpurun_out/driver_scale_n2_single.log
timeout 200 python profiles/tools/driver_scale.py --devices 0,1 > gpurun_out/driver_scale_n2_01.log 2> gpurun_out/driver_scale_n2_01.err; cat gpurun_out/driver_scale_n2_01.log; tail -2 gpurun_out/driver_scale_n2_01.err
timeout 200 python profiles/tools/driver_scale.py --devices 0,0 > gpurun_out/driver_scale_n2_00.log 2> gpurun_out/driver_scale_n2_00.err; cat gpurun_out/driver_scale_n2_00.log
