mkdir -p gpurun_out
rm -f gpurun_out/relerr.log
( BKE_TEST_ERRLOG=gpurun_out/relerr.log timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
  echo "== memcheck"
  timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_r2.log python -m pytest tests/test_gpu_kf.py tests/test_gpu_ukf.py tests/test_gpu_resample.py tests/test_gpu_parity_holes.py -q -m gpu -k "golden or singular or small_shapes or indefinite or sticky or composites or plan_single or rowblock_separate" -x 2>&1 | tail -3
  tail -3 gpurun_out/memcheck_r2.log
  echo "== racecheck"
  timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck_r2.log python -m pytest tests/test_gpu_kf.py tests/test_gpu_ukf.py tests/test_gpu_resample.py -q -m gpu -k "bank_vs_reference_golden or ukf_bank_vs_reference_golden or golden_vectors_from_reference or plan_single" -x 2>&1 | tail -3
  tail -3 gpurun_out/racecheck_r2.log
) > gpurun_out/full2.log 2>&1
cat gpurun_out/full2.log | cut -c1-400
