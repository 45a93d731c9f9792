mkdir -p gpurun_out
SEL='rowblock or batch_filter or bank_vs_reference_golden or user or julier or composite or plan_single or resample_golden or indefinite or in_place or sticky or nones'
( timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2b_memcheck.log python -m pytest tests/test_gpu_kf.py tests/test_gpu_ukf.py tests/test_gpu_parity_holes.py tests/test_gpu_resample.py -q -x -k "$SEL" -p no:cacheprovider 2>&1 | tail -2
  tail -2 gpurun_out/r2b_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2b_racecheck.log python -m pytest tests/test_gpu_kf.py tests/test_gpu_resample.py -q -x -k "rowblock or batch_filter or resample_golden or composite" -p no:cacheprovider 2>&1 | tail -2
  tail -2 gpurun_out/r2b_racecheck.log
) > gpurun_out/run10.log 2>&1
cat gpurun_out/run10.log
