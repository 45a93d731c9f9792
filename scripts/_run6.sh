mkdir -p gpurun_out
( timeout 1100 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
  timeout 300 python scripts/variants_bench.py 2>&1 | grep -i "batch\|C3\|C2" 
  BKE_BATCH_DIRECT=1 timeout 300 python scripts/variants_bench.py 2>&1 | grep -i "batch" | sed 's/^/direct /'
) > gpurun_out/run6.log 2>&1
cat gpurun_out/run6.log
