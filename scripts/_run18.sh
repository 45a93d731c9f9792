mkdir -p gpurun_out
( echo "== tc check"; BKE_KF_TC=2 timeout 300 python scripts/tc_check.py check 2>&1 | tail -1
  echo "== tc time, CTAs per SM sweep"
  for c in 1 2 4 8; do echo "ctas=$c"; BKE_KF_TC_CTAS=$c timeout 200 python scripts/tc_check.py time 2>&1 | grep '"predict ' ; done
  echo "== full gpu tests"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
  echo "== ncu"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_cov_tc -c 1 -s 2 -o gpurun_out/r2c_tc16 -f python scripts/tc_profile.py 16 2>&1 | tail -1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:kf_cov_tc -c 1 -s 2 -o gpurun_out/r2c_tc32 -f python scripts/tc_profile.py 32 2>&1 | tail -1
) > gpurun_out/run18.log 2>&1
cat gpurun_out/run18.log
