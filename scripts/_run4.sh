mkdir -p gpurun_out
export RS_REPS=5
( BKE_RS_IMPL=fused timeout 500 python scripts/rs_fused_check.py quick 2>&1 | grep "^FAIL\|FAILS\|rror" | head -20;
  BKE_RS_IMPL=fused BKE_RS_STAGES=5 BKE_RS_PROF=1 timeout 100 python scripts/rs_onebinade.py 26 2>&1 | grep "ONEBINADE\|RSPROF" | tail -2
  BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 heavy new:8:2:0:5 new:8:2:0:3 new:8:2:0:8 new:4:4:0:6 2>&1 | grep "SWEEP\|RSPROF" | awk '/SWEEP/{print last; print} {last=$0}'
  BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 uniform new:8:2:0:5 2>&1 | grep "SWEEP\|RSPROF" | tail -2
  BKE_RS_IMPL=fused timeout 200 python scripts/rs_trace.py 26 heavy 2>&1 | tail -26
) > gpurun_out/sweep10.log 2>&1
cat gpurun_out/sweep10.log
