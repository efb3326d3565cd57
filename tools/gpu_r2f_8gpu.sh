#!/bin/bash
# r02t (gpurun --gpus 8): N = 8 and N = 4 bench lines on the round-2 final code (cfg2 weak + cfg3 sharded + strong scaling + e2e / PCIe probe), 2-GPU tests
OUT=gpurun_out/r02t; mkdir -p $OUT
{ nvidia-smi topo -m; nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; free -g | head -2; } > $OUT/topo.txt 2>&1
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  tail -c 600 $OUT/bench_n$n.json; tail -3 $OUT/bench_n$n.err
done
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > $OUT/pytest_multigpu.log 2>&1; tail -3 $OUT/pytest_multigpu.log
