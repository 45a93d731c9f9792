mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
( timeout 200 $NCU -k regex:kf42_f32 --launch-skip 3 --launch-count 1 -o gpurun_out/r2b_kf42_diag -f python scripts/kf42_diag_profile.py > /dev/null 2>&1; echo kf42 $?
  timeout 200 $NCU -k regex:k_tile_maps --launch-skip 4 --launch-count 2 -o gpurun_out/r2b_tile_maps -f python scripts/rs_bench.py 26 2 heavy > /dev/null 2>&1; echo maps $?
  timeout 200 $NCU -k regex:k_emit_slow --launch-skip 2 --launch-count 1 -o gpurun_out/r2b_emit_slow -f python scripts/rs_bench.py 26 2 heavy > /dev/null 2>&1; echo emit_slow $?
) > gpurun_out/run12.log 2>&1
cat gpurun_out/run12.log
