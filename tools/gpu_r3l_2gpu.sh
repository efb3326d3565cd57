#!/bin/bash
# r03l (gpurun --gpus 2): the 2-GPU tests, the FIR tests and the N = 2 bench line on the round-2 final tree
OUT=gpurun_out/r03l; mkdir -p $OUT
timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_fir_conv_gpu.py -m gpu -q > $OUT/pytest_multigpu.log 2>&1; tail -3 $OUT/pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
tail -c 600 $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03l/bench_n2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"])
print(json.dumps(d["multi_gpu"], indent=1)[:3500])
PY
