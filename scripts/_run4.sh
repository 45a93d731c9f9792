mkdir -p gpurun_out
export RS_REPS=5
( timeout 400 python scripts/rs_fused_check.py quick 2>&1 | grep "^FAIL\|FAILS\|rror" | head -20;
  BKE_RS_STAGES=3 BKE_RS_PROF=1 timeout 100 python scripts/rs_onebinade.py 26 2>&1 | grep "ONEBINADE\|RSPROF" | tail -2
  BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 heavy new:8:2:0:3 new:8:2:0:2 new:8:2:0:5 new:4:4:0:3 2>&1 | grep "SWEEP\|RSPROF" | awk '/SWEEP/{print last; print} {last=$0}'
  BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 uniform new:8:2:0:3 2>&1 | grep "SWEEP\|RSPROF" | tail -2
) > gpurun_out/sweep8.log 2>&1
cat gpurun_out/sweep8.log
