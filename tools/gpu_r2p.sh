#!/bin/bash
# r02p: validation of the round-2 final code: full GPU suite, smoke, bench + reference arm, ncu launch list of the bench command, compute-sanitizer
OUT=gpurun_out/r02p; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extras --no-multi --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1; wc -l $OUT/launches.csv
bash tools/gpu_sanitize.sh > $OUT/sanitize.log 2>&1; tail -6 $OUT/sanitize.log; cp gpurun_out/sanitize/*.log $OUT/ 2>/dev/null
