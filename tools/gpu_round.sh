#!/bin/bash
# One gpurun call: GPU tests, bench line, ncu launch list, ncu --set full of the hot kernels.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 1500 $OUT/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
for a in "8 600 1024 256" "128 60 4096 1024" "8 600 2048 512" "8 600 512 128"; do timeout 120 python tools/run_stft.py $a 10; done > $OUT/stft_shapes.txt 2>&1
timeout 120 python tools/run_istft.py 32 60 1024 256 10 > $OUT/istft.txt 2>&1
timeout 120 python tools/run_istft.py 8 60 4096 1024 5 >> $OUT/istft.txt 2>&1
timeout 200 python tools/run_fir.py 64 600 2049 5 > $OUT/fir.txt 2>&1
timeout 200 python tools/run_fir.py 64 600 255 5 >> $OUT/fir.txt 2>&1
cat $OUT/stft_shapes.txt $OUT/istft.txt $OUT/fir.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft_full -f python tools/run_stft.py 8 600 1024 256 2 > $OUT/ncu_stft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_kernel -s 2 -c 1 -o $OUT/istft_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu_istft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_ols -s 2 -c 1 -o $OUT/fir_full -f python tools/run_fir.py 64 60 2049 2 > $OUT/ncu_fir.log 2>&1
ls -la $OUT
