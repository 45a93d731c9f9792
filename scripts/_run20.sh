mkdir -p gpurun_out
( echo "== MN variant: check"; BKE_KF_TC_MN=1 BKE_KF_TC=2 timeout 300 python scripts/tc_check.py check 2>&1 | tail -3
  echo "== MN variant: time"; BKE_KF_TC_MN=1 timeout 200 python scripts/tc_check.py time 2>&1 | grep '"predict '
  echo "== K-major (default): time"; timeout 200 python scripts/tc_check.py time 2>&1 | grep '"predict '
  echo "== MN variant: tests"; BKE_KF_TC_MN=1 timeout 600 python -m pytest tests/test_gpu_kf_tc.py -q 2>&1 | tail -2
  echo "== memcheck"
  timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2c_memcheck.log python -m pytest tests/test_gpu_kf_tc.py tests/test_gpu_next_rows.py tests/test_gpu_kf.py tests/test_torch_ops.py -q -x -p no:cacheprovider -k "(tc_predict and not 40003) or tc_fused or residual_resample_vs_reference_golden or residual_cumsum or (small_shapes and (16 or 32)) or twin_kf" 2>&1 | tail -2
  tail -3 gpurun_out/r2c_memcheck.log
  echo "== racecheck"
  timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2c_racecheck.log python -m pytest tests/test_gpu_kf_tc.py tests/test_gpu_next_rows.py -q -x -p no:cacheprovider -k "(tc_predict and 1037) or tc_fused or residual_cumsum" 2>&1 | tail -2
  tail -3 gpurun_out/r2c_racecheck.log
) > gpurun_out/run20.log 2>&1
cat gpurun_out/run20.log
