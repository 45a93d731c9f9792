mkdir -p gpurun_out
( echo "== tc check"; BKE_KF_TC=2 timeout 300 python scripts/tc_check.py check 2>&1 | tail -2
  echo "== tc time (2 = always, 1 = default policy, 0 = off)"
  for m in 2 1 0; do BKE_KF_TC=$m timeout 200 python scripts/tc_check.py time 2>&1 | tail -4; done
  echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_next_rows.py -x -q -k "residual" 2>&1 | tail -3
  timeout 600 python -m pytest tests/test_gpu_kf_tc.py tests/test_torch_ops.py -x -q 2>&1 | tail -5
  BKE_KF_TC=2 timeout 600 python -m pytest tests/test_gpu_kf_tc.py -x -q 2>&1 | tail -3
  echo "== ncu"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_cov_tc -c 1 -s 2 -o gpurun_out/r2c_tc16 -f python scripts/tc_profile.py 16 2>&1 | tail -2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_cov_tc -c 1 -s 2 -o gpurun_out/r2c_tc32 -f python scripts/tc_profile.py 32 2>&1 | tail -2
) > gpurun_out/run16.log 2>&1
cat gpurun_out/run16.log
