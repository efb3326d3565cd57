#!/bin/bash
# One gpurun call: GPU tests, bench line, reference arm, kernel timings, ncu launch list, ncu --set full of the hot kernels.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt; free -g >> $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 3000 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
{ for a in "8 600 1024 256" "128 60 4096 1024" "8 600 2048 512" "8 600 512 128" "32 60 8192 2048"; do timeout 120 python tools/run_stft.py $a 10; done
for a in "8 600 1024 250" "8 600 1024 441"; do timeout 120 python tools/run_stft.py $a 10; done
for a in "32 60 1024 256" "32 60 512 128" "32 60 2048 512" "32 60 4096 1024" "32 60 1024 512" "32 60 1024 128" "32 60 1024 250"; do timeout 120 python tools/run_istft.py $a 10; done
for a in "32 60 1024 256" "32 60 1024 512" "32 60 512 128" "32 60 2048 512" "32 60 4096 1024"; do timeout 120 python tools/run_istft_c2r.py $a 10; done
timeout 200 python tools/run_fir.py 64 600 2049 5; timeout 200 python tools/run_fir.py 64 600 255 5; timeout 200 python tools/run_fir.py 64 600 65 5
timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 48000; timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 16000; } > $OUT/kernel_timings.txt 2>&1
cat $OUT/kernel_timings.txt
timeout 200 python tools/run_e2e.py 8 600 3 > $OUT/e2e.txt 2>&1; cat $OUT/e2e.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft_full -f python tools/run_stft.py 8 600 1024 256 2 > $OUT/ncu_stft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola -s 2 -c 1 -o $OUT/istft_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu_istft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_ols_pg -s 2 -c 1 -o $OUT/fir_full -f python tools/run_fir.py 64 60 2049 2 > $OUT/ncu_fir.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola_c2r -s 2 -c 1 -o $OUT/istft_c2r_full -f python tools/run_istft_c2r.py 32 60 1024 256 2 > $OUT/ncu_c2r.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 1 -c 1 -o $OUT/stft_mel_full -f python tools/run_mel.py 8 600 1024 256 128 0 48000 > $OUT/ncu_stft_mel.log 2>&1
ls -la $OUT
