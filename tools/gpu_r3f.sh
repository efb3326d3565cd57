#!/bin/bash
# r03f: bench + reference arm, ncu launch list of the bench command, ncu --set full of the three kernels that went to packed fp32x2
OUT=gpurun_out/r03f; mkdir -p $OUT
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extras --no-multi --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1; wc -l $OUT/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola -s 2 -c 1 -o $OUT/istft_pk_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu_istft.log 2>&1; tail -1 $OUT/ncu_istft.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_ols_r2c -s 2 -c 1 -o $OUT/fir_pk_full -f python tools/run_fir.py 64 60 2049 2 > $OUT/ncu_fir.log 2>&1; tail -1 $OUT/ncu_fir.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft4096_pk_full -f python tools/run_stft.py 128 60 4096 1024 2 > $OUT/ncu_stft4096.log 2>&1; tail -1 $OUT/ncu_stft4096.log
du -sh gpurun_out
