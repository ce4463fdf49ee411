# after profiles/tools/gpu_final.sh came back: turn gpurun_out/ into the committed round-2 evidence
set -e
cd "$(dirname "$0")/../.."
H=$(python -c "import sys; sys.path.insert(0,'.'); from metsim_b200 import build; print(build.source_hash())")
grep -q "$H" gpurun_out/smoke.log || { echo "gpurun_out/ is not from the present sources ($H)"; exit 1; }
mkdir -p profiles/r02
cp gpurun_out/bench.log profiles/r02/bench_n1.json
cp gpurun_out/bench_reference.log profiles/r02/bench_reference.json
cp gpurun_out/launches.csv profiles/r02/launch_list_bench.csv
python profiles/tools/launch_sum.py gpurun_out/launches.csv > profiles/r02_launch_list_summary.txt
python profiles/tools/ncu_summary.py gpurun_out/r02_step_kernels.ncu-rep > profiles/r02_step_kernels_ncu_full.txt 2>&1
grep -q "$H" gpurun_out/disagg_traffic.json && cp gpurun_out/disagg_traffic.json profiles/disagg_traffic.json
cp gpurun_out/bench_c2.log profiles/r02/bench_c2_global_half_deg.json
grep -v "Warn\|_warn_once" gpurun_out/timeline_c3.log > profiles/r02_timeline_conus.txt
grep -v "Warn\|_warn_once" gpurun_out/timeline_c2.log > profiles/r02_timeline_global_half_deg.txt
echo "evidence collected for sources $H"
