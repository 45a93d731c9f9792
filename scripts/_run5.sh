mkdir -p gpurun_out
( timeout 400 python scripts/rs_fused_check.py quick 2>&1 | grep -v "^ok" | tail -8
  for k in heavy uniform; do timeout 120 python scripts/rs_bench.py 26 10 $k 2>&1 | tail -1; done
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_front.csv python scripts/rs_bench.py 26 2 heavy > /dev/null 2>&1
  tail -8 gpurun_out/r2_launches_front.csv | cut -d, -f5,15-
) > gpurun_out/run5.log 2>&1
cat gpurun_out/run5.log
