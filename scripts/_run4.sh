mkdir -p gpurun_out
export RS_REPS=10
( for v in 0 1; do BKE_RS_E2=$v timeout 120 python scripts/rs_sweep.py 26 heavy old:8:0:0 2>&1 | grep SWEEP | sed "s/^/e2=$v /"; BKE_RS_E2=$v timeout 120 python scripts/rs_sweep.py 26 uniform old:8:0:0 2>&1 | grep SWEEP | sed "s/^/e2=$v /"; done
  BKE_RS_E2=1 timeout 600 python -m pytest tests/test_gpu_resample.py -x -q 2>&1 | tail -2
) > gpurun_out/sweep12.log 2>&1
cat gpurun_out/sweep12.log
