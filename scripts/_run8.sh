mkdir -p gpurun_out
( for r in 1 2; do
  for k in heavy uniform; do
    BKE_LIB_PATH=$PWD/filterpy_b200/_C/libbke_prev.so timeout 120 python scripts/rs_bench.py 26 10 $k 2>&1 | tail -1 | sed 's/^/prev /'
    timeout 120 python scripts/rs_bench.py 26 10 $k 2>&1 | tail -1 | sed 's/^/new  /'
  done; done
  timeout 600 python -m pytest tests/test_gpu_resample.py tests/test_gpu_ukf.py -x -q 2>&1 | tail -3
) > gpurun_out/run8.log 2>&1
cat gpurun_out/run8.log
