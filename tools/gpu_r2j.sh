#!/bin/bash
# r02j: validation of the round-2 state: full GPU suite, smoke, timed bench + reference arm, ncu launch list of the bench command, ncu --set full of the hot kernels
OUT=gpurun_out/r02j; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
/usr/bin/time -v -o $OUT/bench_time.txt timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; grep -E "Elapsed|Maximum resident" $OUT/bench_time.txt; tail -c 500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
/usr/bin/time -v -o $OUT/ref_time.txt timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; grep -E "Elapsed" $OUT/ref_time.txt; cut -c1-200 $OUT/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extras --no-multi --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1; wc -l $OUT/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft1024_full -f python tools/run_stft.py 8 600 1024 256 2 > $OUT/ncu1.log 2>&1; tail -1 $OUT/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft4096_full -f python tools/run_stft.py 128 60 4096 1024 2 > $OUT/ncu2.log 2>&1; tail -1 $OUT/ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola -s 2 -c 1 -o $OUT/istft_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu3.log 2>&1; tail -1 $OUT/ncu3.log
