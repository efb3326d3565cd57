# end-of-round confirmation on the final code: full GPU suite, bench + reference arm, log-mel timings
TAG=${1:-r01T}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -2 $OUT/bench_n1.err; wc -l $OUT/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
{ timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 48000; timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 16000; } > $OUT/mel.txt 2>&1; cat $OUT/mel.txt
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
