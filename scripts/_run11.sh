mkdir -p gpurun_out
( nvidia-smi -L | head -3
  timeout 600 python -m pytest tests/test_gpu_resample.py -x -q -k "nccl or sharded or plan" 2>&1 | tail -3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/dist_resample_check.py 26 plan 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r2b_bench_2gpu.json
  python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench_2gpu.json").read())
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
for k,v in d.items():
    if isinstance(v,dict) and "roofline" in v and k!="roofline" and "ms_per_step" in v: print(k, round(v["ms_per_step"],4), round(v["roofline"]["frac"],3))
r=d["resample"]; print("resample", r["ms"], r["value"], r["bit_exact_vs_oracle"], r.get("exchange"))
PY
) > gpurun_out/run11.log 2>&1
cat gpurun_out/run11.log
