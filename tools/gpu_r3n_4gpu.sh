#!/bin/bash
# r03n (gpurun --gpus 4): N = 4 bench line (device legs only: headline weak scaling, cfg3 sharded, cfg2 strong scaling; the e2e legs are unchanged since r02t)
OUT=gpurun_out/r03n; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras > $OUT/bench_n4.json 2> $OUT/bench_n4.err
tail -3 $OUT/bench_n4.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03n/bench_n4.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"])
print(json.dumps(d["multi_gpu"], indent=1)[:3500])
PY
