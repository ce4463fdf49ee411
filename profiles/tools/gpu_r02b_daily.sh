# day-chunk length of the daily pass (run-time knob MSB_DAILY_CHUNK on the product build)
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for c in 0 37 61 73 92 122 183 365; do
  env MSB_DAILY_CHUNK=$c timeout 120 python profiles/tools/timeline.py --steps 2 > gpurun_out/variant_daily$c.log 2>&1
  echo "== chunk $c"; grep -h "events, prelude threads all\|daily_fused" gpurun_out/variant_daily$c.log | head -3
done
