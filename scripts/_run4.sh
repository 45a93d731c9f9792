mkdir -p gpurun_out
export RS_REPS=10
( timeout 900 python -m pytest tests/test_gpu_resample.py -x -q 2>&1 | tail -3
  timeout 120 python scripts/rs_sweep.py 26 heavy old:8:0:0 2>&1 | grep SWEEP
  BKE_RS_EMIT=1 timeout 120 python scripts/rs_sweep.py 26 heavy old:8:0:0 2>&1 | grep SWEEP | sed "s/^/emit1 /"
  timeout 120 python scripts/rs_sweep.py 26 uniform old:8:0:0 2>&1 | grep SWEEP
  BKE_RS_IMPL=fused timeout 400 python scripts/rs_fused_check.py quick 2>&1 | grep "^FAIL\|FAILS\|rror" | head
) > gpurun_out/sweep9.log 2>&1
cat gpurun_out/sweep9.log
