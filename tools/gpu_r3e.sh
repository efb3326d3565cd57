#!/bin/bash
# r03e: validation of the final round-2 tree: full GPU suite, smoke, bench + reference arm, ncu launch list of the bench command,
# ncu --set full of the kernels that changed (packed fp32x2 plans), kernel timings over shapes, compute-sanitizer
OUT=gpurun_out/r03e; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
{ for a in "8 600 1024 256" "128 60 4096 1024" "8 600 2048 512" "8 600 512 128" "32 60 8192 2048" "8 600 1024 250"; do timeout 120 python tools/run_stft.py $a 10; done
for a in "32 60 1024 256" "32 60 512 128" "32 60 2048 512" "32 60 4096 1024" "32 60 1024 512" "32 60 1024 128" "32 60 1024 250"; do timeout 120 python tools/run_istft.py $a 10; done
for a in "32 60 1024 256" "32 60 1024 512" "32 60 2048 512"; do timeout 120 python tools/run_istft_c2r.py $a 10; done
for k in 2049 1025 4097 10001 513 255 65; do timeout 200 python tools/run_fir.py 64 600 $k 3; done; } > $OUT/kernel_timings.txt 2>&1; cat $OUT/kernel_timings.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extras --no-multi --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1; wc -l $OUT/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola -s 2 -c 1 -o $OUT/istft_pk_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu_istft.log 2>&1; tail -1 $OUT/ncu_istft.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_ols_r2c -s 2 -c 1 -o $OUT/fir_pk_full -f python tools/run_fir.py 64 60 2049 2 > $OUT/ncu_fir.log 2>&1; tail -1 $OUT/ncu_fir.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft4096_pk_full -f python tools/run_stft.py 128 60 4096 1024 2 > $OUT/ncu_stft4096.log 2>&1; tail -1 $OUT/ncu_stft4096.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft1024_full -f python tools/run_stft.py 8 600 1024 256 2 > $OUT/ncu_stft1024.log 2>&1; tail -1 $OUT/ncu_stft1024.log
bash tools/gpu_sanitize.sh > $OUT/sanitize.log 2>&1; tail -6 $OUT/sanitize.log; cp gpurun_out/sanitize/*.log $OUT/ 2>/dev/null
ls -la $OUT | head -40
