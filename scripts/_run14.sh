mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_next_rows.py -x -q -k "residual or multinomial_public" 2>&1 | tail -5
  timeout 900 python -m pytest tests/test_gpu_kf.py -x -q -k "small_shapes or rowblock" 2>&1 | tail -5
  timeout 600 python scripts/r2c_bench.py 2>&1 | tail -20
) > gpurun_out/run14.log 2>&1
cat gpurun_out/run14.log
( for v in 0 1 0 1; do BKE_RS_E2=$v timeout 120 python scripts/rs_bench.py 26 10 heavy 2>&1 | tail -1 | cut -c1-160 | sed "s/^/E2=$v /"; done ) >> gpurun_out/run14.log 2>&1
tail -4 gpurun_out/run14.log
