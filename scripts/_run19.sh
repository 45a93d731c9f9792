mkdir -p gpurun_out
( echo "== tc check"; BKE_KF_TC=2 timeout 300 python scripts/tc_check.py check 2>&1 | tail -1
  echo "== tc time"; for m in 1 0; do BKE_KF_TC=$m timeout 200 python scripts/tc_check.py time 2>&1 | tail -4; done
  echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kf.py tests/test_gpu_kf_tc.py -q 2>&1 | tail -3
  echo "== bench"; timeout 900 python bench.py > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_err.log; python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_1gpu.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','roofline','frac')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','kf_c3','ukf_c4','resample','kf_tc_predict_16','kf_tc_predict_32','kf_c2_diagnostics','kf_batch_filter','clocks','gpu_launches')})"
) > gpurun_out/run19.log 2>&1
cat gpurun_out/run19.log
