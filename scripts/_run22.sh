mkdir -p gpurun_out
rm -f gpurun_out/errlog.txt
( echo "== tests with the error log"; BKE_TEST_ERRLOG=gpurun_out/errlog.txt timeout 900 python -m pytest tests/test_gpu_kf.py tests/test_gpu_kf_tc.py -q 2>&1 | tail -6
  grep -E "oracle|16" gpurun_out/errlog.txt | sort -t= -k2 -g | tail -12
  echo "== bench"; timeout 900 python bench.py > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log; python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_1gpu.json').read().strip().splitlines()[-1])
for k in ('kf_tc_predict_16','kf_tc_predict_32','kf_tc_step_16','kf_tc_step_32','resample'): print(k, json.dumps(d.get(k))[:300])
print('value', d['value'], d['roofline']['frac'], d['clocks'])"
) > gpurun_out/run22.log 2>&1
cat gpurun_out/run22.log
