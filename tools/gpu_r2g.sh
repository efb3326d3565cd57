#!/bin/bash
# r02g: full GPU suite on the current code, bench line, 1-channel shapes (strong scaling by channel), compute-sanitizer
OUT=gpurun_out/r02g; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 700 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
{ timeout 120 python tools/run_stft.py 1 600 1024 256 20; timeout 120 python tools/run_stft.py 2 600 1024 256 20; timeout 120 python tools/run_stft.py 8 75 1024 256 20; timeout 120 python tools/run_stft.py 8 150 1024 256 20
timeout 120 python tools/run_istft.py 32 60 1024 256 10; timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
bash tools/gpu_sanitize.sh > $OUT/sanitize.log 2>&1; tail -12 $OUT/sanitize.log; cp gpurun_out/sanitize/*.log $OUT/ 2>/dev/null
