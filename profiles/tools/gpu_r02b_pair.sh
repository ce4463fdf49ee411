# A/B of the step-pair variant of thermo_knot_kernel (profiles/tools/experiments/disagg_interior_pair.cu) at
# 6 / 5 / 4 resident CTAs per SM against the product build, parity of one variant, driver re-run
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in base pair5 pair6 pair4; do
  MSB_LIBRARY=$GRAFT_REPO_ROOT/profiles/variants/lib_$v.so timeout 200 python profiles/tools/timeline.py --steps 2 > gpurun_out/variant_$v.log 2>&1
  echo "== $v"; grep -h "events, prelude threads all\|thermo_knot" gpurun_out/variant_$v.log | head -6
done
( time MSB_LIBRARY=$GRAFT_REPO_ROOT/profiles/variants/lib_pair5.so timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_pair5.log 2>&1; echo "pytest pair5 rc=$?" >> gpurun_out/pytest_pair5.log
tail -6 gpurun_out/pytest_pair5.log
timeout 200 python profiles/tools/driver_scale.py > gpurun_out/driver_scale2.log 2> gpurun_out/driver_scale2.err; cat gpurun_out/driver_scale2.log
