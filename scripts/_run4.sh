mkdir -p gpurun_out
export RS_REPS=5
( timeout 300 python scripts/rs_fused_check.py quick 2>&1 | grep -c "^ok"; 
  for lbk in 1 4; do BKE_RS_LBK=$lbk BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 heavy new:8:2:0:2 2>&1 | grep "SWEEP\|RSPROF" | tail -2 | sed "s/^/lbk=$lbk /"; done
  BKE_RS_LBK=4 BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 heavy new:8:1:0:3 new:4:3:0:3 new:4:2:0:4 new:4:4:0:2 2>&1 | grep "SWEEP\|RSPROF" | awk '/SWEEP/{print last; print} {last=$0}'
  BKE_RS_LBK=4 BKE_RS_PROF=1 timeout 120 python scripts/rs_sweep.py 26 uniform new:8:2:0:2 2>&1 | grep "SWEEP\|RSPROF" | tail -2
) > gpurun_out/sweep5.log 2>&1
cat gpurun_out/sweep5.log
