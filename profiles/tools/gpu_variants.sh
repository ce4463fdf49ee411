# A/B of differently-compiled builds (profiles/tools/build_variants.py) on the bench step: per-kernel times from the activity timeline
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in base ov0 ov6 ov5 ov4 cap5 prec8; do
  MSB_LIBRARY=$GRAFT_REPO_ROOT/profiles/variants/lib_$v.so timeout 300 python profiles/tools/timeline.py --steps 2 > gpurun_out/variant_$v.log 2>&1
  echo "== $v"; grep -h "events, prelude threads all\|timeline of\|sw_day_kernel\|prec_day_kernel\|thermo_knot" gpurun_out/variant_$v.log
done
