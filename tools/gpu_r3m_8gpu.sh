#!/bin/bash
# r03m (gpurun --gpus 8): N = 8 bench line (device legs only: headline weak scaling, cfg3 sharded, cfg2 strong scaling; the e2e legs are unchanged since r02t)
OUT=gpurun_out/r03m; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras > $OUT/bench_n8.json 2> $OUT/bench_n8.err
tail -3 $OUT/bench_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03m/bench_n8.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"])
print(json.dumps(d["multi_gpu"], indent=1)[:3500])
PY
