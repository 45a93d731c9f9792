mkdir -p gpurun_out
export RS_REPS=5
( timeout 300 python scripts/rs_fused_check.py quick 2>&1 | grep "^FAIL\|FAILS\|rror" | head -20;
  BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 heavy new:8:2:0:2 new:8:1:0:3 new:4:3:0:3 new:4:2:0:4 2>&1 | grep "SWEEP\|RSPROF" | awk '/SWEEP/{print last; print} {last=$0}'
  BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 uniform new:8:2:0:2 2>&1 | grep "SWEEP\|RSPROF" | tail -2
  timeout 200 python scripts/rs_trace.py 26 heavy 2>&1 | tail -28
) > gpurun_out/sweep6.log 2>&1
cat gpurun_out/sweep6.log
