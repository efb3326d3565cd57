#!/bin/bash
# r02d: full GPU test suite with the new host pipeline / whole-channel parity tests, bench line, compute-sanitizer on the new kernels
OUT=gpurun_out/r02d; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 1500 $OUT/bench_n1.json; tail -5 $OUT/bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json | cut -c1-600
