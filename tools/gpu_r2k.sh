#!/bin/bash
# r02k: timed bench line + reference arm (wall clock of each command) on the round-2 final code
OUT=gpurun_out/r02k; mkdir -p $OUT
s=$(date +%s.%N); timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; e=$(date +%s.%N); echo "bench.py wall seconds: $(echo "$e - $s" | bc)" | tee $OUT/wall.txt; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
s=$(date +%s.%N); timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; e=$(date +%s.%N); echo "bench.py --impl reference wall seconds: $(echo "$e - $s" | bc)" | tee -a $OUT/wall.txt; cut -c1-200 $OUT/bench_ref.json
