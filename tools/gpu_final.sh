# final checks of the round: FIR tests + short-filter timings, full bench + reference arm
TAG=${1:-r01P}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_fir_conv_gpu.py -x -q > $OUT/pytest_fir.log 2>&1; tail -4 $OUT/pytest_fir.log
{ for k in 17 33 65 129 255 2049; do timeout 200 python tools/run_fir.py 64 600 $k 5; done; NXS_FIR_VARIANT=2 timeout 200 python tools/run_fir.py 64 600 65 5; } > $OUT/fir_timings.txt 2>&1; cat $OUT/fir_timings.txt
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -2 $OUT/bench_n1.err; wc -l $OUT/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
