# A/B: partial mixes (thermo + shortwave, thermo + precipitation in one launch), store policies of thermo_knot_kernel,
# tile lengths of the interior kernels (run-time knobs on the product build)
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() {  # name, library, extra env
  env MSB_LIBRARY=$GRAFT_REPO_ROOT/profiles/variants/lib_$2.so $3 timeout 120 python profiles/tools/timeline.py --steps 2 > gpurun_out/variant_$1.log 2>&1
  echo "== $1"; grep -h "events, prelude threads all\|thermo_knot\|interior_mixed\|sw_day\|prec_day" gpurun_out/variant_$1.log | head -5
}
run base base ""
run mixts6 mixts6 ""
run mixts7 mixts7 ""
run mixtp6 mixtp6 ""
run plainst plainst ""
run stwt stwt ""
run stcg stcg ""
run knot8 base MSB_KNOT_TILE=8
run knot32 base MSB_KNOT_TILE=32
run day8 base MSB_DAY_TILE=8
run day32 base MSB_DAY_TILE=32
run base_again base ""
