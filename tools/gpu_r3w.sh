#!/bin/bash
# r03w: validation of the final tree (ISTFT nfft 128 on the register-overlap-add kernel, sub-warp groups synchronise over their own lanes) + compute-sanitizer: full GPU suite, smoke, bench + reference arm
OUT=gpurun_out/r03w; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
{ timeout 120 python tools/run_stft.py 8 600 512 128 10; timeout 120 python tools/run_stft.py 8 600 256 64 10; timeout 120 python tools/run_stft.py 8 600 128 32 10; timeout 120 python tools/run_istft.py 32 60 1024 250 10; timeout 120 python tools/run_istft.py 32 60 1024 441 10; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
timeout 120 python tools/run_istft.py 64 60 128 32 10 >> $OUT/timings.txt 2>&1
bash tools/gpu_sanitize.sh > $OUT/sanitize.log 2>&1; tail -8 $OUT/sanitize.log; cp gpurun_out/sanitize/*.log $OUT/ 2>/dev/null
