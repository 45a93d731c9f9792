mkdir -p gpurun_out
( for r in 1 2; do
    BKE_LIB_PATH=$PWD/filterpy_b200/_C/libbke_prev.so timeout 300 python scripts/variants_bench.py 2>&1 | grep "C2 kf" | cut -c1-140 | sed 's/^/prev /'
    timeout 300 python scripts/variants_bench.py 2>&1 | grep "C2 kf" | cut -c1-140 | sed 's/^/new  /'
  done
  timeout 300 python -m pytest tests/test_gpu_kf.py tests/test_gpu_parity_holes.py -x -q 2>&1 | tail -2
) > gpurun_out/run9.log 2>&1
cat gpurun_out/run9.log
