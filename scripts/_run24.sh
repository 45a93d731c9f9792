mkdir -p gpurun_out
( timeout 600 python bench.py > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log; python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_1gpu.json').read().strip().splitlines()[-1])
for k in ('kf_tc_predict_16','kf_tc_predict_32','kf_tc_step_16','kf_tc_step_32'): print(k, d[k]['ms_per_step'], d[k]['roofline']['frac'])
print('value', d['value'], d['roofline']['frac'], d['clocks'], d['resample']['ms'])"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:kf_cov_tc -c 1 -s 2 -o gpurun_out/r2c_tc16_step -f python scripts/tc_profile.py 16 step 2>&1 | tail -1
) > gpurun_out/run24.log 2>&1
cat gpurun_out/run24.log
