mkdir -p gpurun_out
( BKE_RS_MAPS=1 BKE_RS_SCAN=1 timeout 400 python scripts/rs_fused_check.py quick 2>&1 | grep -v "^ok" | tail -4
  for r in 1 2; do
   for cfg in "0 0" "1 0" "0 1" "1 1"; do set -- $cfg
    BKE_RS_MAPS=$1 BKE_RS_SCAN=$2 timeout 120 python scripts/rs_bench.py 26 10 heavy 2>&1 | tail -1 | cut -c1-130 | sed "s/^/maps=$1 scan=$2 /"
   done
  done
  BKE_RS_MAPS=1 BKE_RS_SCAN=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches_resample.csv python scripts/rs_bench.py 26 2 heavy > /dev/null 2>&1
  tail -10 gpurun_out/r2b_launches_resample.csv | cut -d, -f5,15-
) > gpurun_out/run13.log 2>&1
cat gpurun_out/run13.log
