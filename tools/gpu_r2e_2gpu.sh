#!/bin/bash
# r02q (gpurun --gpus 2): the 2-GPU tests and the N = 2 bench line on the round-2 final code
OUT=gpurun_out/r02q; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > $OUT/pytest_multigpu.log 2>&1; tail -5 $OUT/pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
tail -c 3000 $OUT/bench_n2.json; tail -15 $OUT/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; cut -c1-300 $OUT/bench_ref_n2.json
