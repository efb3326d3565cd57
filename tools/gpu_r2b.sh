#!/bin/bash
# r02b: paired split pass + new host pipeline: parity tests, old-vs-new timings, bench line
OUT=gpurun_out/r02b; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
{ for v in 4 0; do echo "NXS_STFT_VARIANT=$v (4 = unpaired 4096)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 128 60 4096 1024 10; done
for v in 3 0; do echo "NXS_STFT_VARIANT=$v (3 = unpaired 2048)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 2048 512 10; done
timeout 120 python tools/run_stft.py 8 600 1024 256 10
timeout 120 python tools/run_stft.py 32 60 8192 2048 10; } > $OUT/timings.txt 2>&1
cat $OUT/timings.txt
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 6000 $OUT/bench_n1.json; tail -5 $OUT/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft4096_full -f python tools/run_stft.py 128 60 4096 1024 2 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
