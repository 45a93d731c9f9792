mkdir -p gpurun_out
( echo "== tc check"; BKE_KF_TC=2 timeout 300 python scripts/tc_check.py check 2>&1 | tail -1
  echo "== time: fused in the tile (default)"; timeout 200 python scripts/tc_check.py time 2>&1 | tail -4
  echo "== time: BKE_KF_TC_FUSED=0"; BKE_KF_TC_FUSED=0 timeout 200 python scripts/tc_check.py time 2>&1 | grep "update"
  echo "== time: BKE_KF_TC=0"; BKE_KF_TC=0 timeout 200 python scripts/tc_check.py time 2>&1 | grep "update"
  echo "== tc tests"; timeout 600 python -m pytest tests/test_gpu_kf_tc.py -q 2>&1 | tail -8
  echo "== full gpu tests"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
) > gpurun_out/run21.log 2>&1
cat gpurun_out/run21.log
