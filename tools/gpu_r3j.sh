#!/bin/bash
# r03j: validation of the final round-2 tree: full GPU suite, smoke, bench + reference arm; SM clock / power sampled while the 106 GB cfg3 call runs
OUT=gpurun_out/r03j; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 100 > $OUT/smi_during_cfg3_1024ch.txt &
SMI=$!
timeout 300 python tools/run_stft.py 1024 60 4096 1024 40 > $OUT/cfg3_1024ch.txt 2>&1; cat $OUT/cfg3_1024ch.txt
NXS_STFT_VARIANT=9 timeout 300 python tools/run_stft.py 1024 60 4096 1024 40 >> $OUT/cfg3_1024ch.txt 2>&1; tail -1 $OUT/cfg3_1024ch.txt
kill $SMI; sort $OUT/smi_during_cfg3_1024ch.txt | uniq -c | sort -rn | head -12
