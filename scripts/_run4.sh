mkdir -p gpurun_out
export RS_REPS=10
( timeout 900 python -m pytest tests/test_gpu_resample.py -x -q 2>&1 | tail -2
  timeout 120 python scripts/rs_sweep.py 26 heavy old:8:0:0 2>&1 | grep SWEEP
  timeout 120 python scripts/rs_sweep.py 26 uniform old:8:0:0 2>&1 | grep SWEEP
  timeout 120 python scripts/rs_sweep.py 24 heavy old:8:0:0 2>&1 | grep SWEEP
) > gpurun_out/sweep11.log 2>&1
cat gpurun_out/sweep11.log
