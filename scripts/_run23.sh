mkdir -p gpurun_out
rm -f gpurun_out/errlog.txt
( echo "== tc check"; BKE_KF_TC=2 timeout 300 python scripts/tc_check.py check 2>&1 | tail -1
  echo "== time"; timeout 200 python scripts/tc_check.py time 2>&1 | tail -4
  echo "== full gpu tests"; BKE_TEST_ERRLOG=gpurun_out/errlog.txt timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
  grep "oracle" gpurun_out/errlog.txt | sort | uniq | tail -8
) > gpurun_out/run23.log 2>&1
cat gpurun_out/run23.log
