mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
( timeout 600 python -m pytest tests/test_gpu_ukf.py -x -q 2>&1 | tail -3
  timeout 200 $NCU -k regex:ukf_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r2_ukf_f64 -f python scripts/ukf_profile.py > /dev/null 2>&1; echo ukf $?
  timeout 200 $NCU -k regex:kf_rowblock --launch-skip 2 --launch-count 1 -o gpurun_out/r2_rb_diag -f python scripts/rb_profile.py diag > /dev/null 2>&1; echo rb $?
  timeout 200 $NCU -k regex:kf_batch --launch-skip 2 --launch-count 1 -o gpurun_out/r2_batch -f python scripts/batch_profile.py > /dev/null 2>&1; echo batch $?
  timeout 200 $NCU -k regex:k_chain_cluster --launch-skip 2 --launch-count 1 -o gpurun_out/r2_chain_cluster -f python scripts/rs_bench.py 26 2 heavy > /dev/null 2>&1; echo chain $?
  timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2b.json; head -c 600 gpurun_out/bench_r2b.json; echo
  ls -la gpurun_out/*.ncu-rep
) > gpurun_out/run7.log 2>&1
cat gpurun_out/run7.log
