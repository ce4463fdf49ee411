# A/B of the mixed-role interior kernel (profiles/tools/experiments/make_mixed.py) against the product build
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in base mixed6 mixed7; do
  MSB_LIBRARY=$GRAFT_REPO_ROOT/profiles/variants/lib_$v.so timeout 200 python profiles/tools/timeline.py --steps 2 > gpurun_out/variant_$v.log 2>&1
  echo "== $v"; grep -h "events, prelude threads all\|thermo_knot\|interior_mixed\|sw_day\|prec_day" gpurun_out/variant_$v.log | head -6
done
