mkdir -p gpurun_out
( echo "== tc check"; timeout 300 python scripts/tc_check.py check 2>&1 | tail -30
  echo "== tc time"; timeout 200 python scripts/tc_check.py time 2>&1 | tail -6
  BKE_KF_TC=0 timeout 200 python scripts/tc_check.py time 2>&1 | tail -6
  echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_next_rows.py -x -q -k "residual" 2>&1 | tail -5
  timeout 600 python -m pytest tests/test_gpu_kf.py -x -q -k "small_shapes or rowblock" 2>&1 | tail -5
  echo "== bench"; timeout 600 python scripts/r2c_bench.py 2>&1 | tail -20
) > gpurun_out/run15.log 2>&1
cat gpurun_out/run15.log
