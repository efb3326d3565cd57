#!/bin/bash
# r03y: large STFT calls walked in channel blocks: parity (same bits), timing over footprints, then the full suite + bench on the final tree
OUT=gpurun_out/r03y; mkdir -p $OUT
timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_full_size_gpu.py tests/test_host_pipeline_gpu.py -m gpu -q > $OUT/pytest_stft.log 2>&1; echo "stft + full size + host: $(tail -1 $OUT/pytest_stft.log)"
{ for c in 256 1024; do timeout 200 python tools/run_stft.py $c 60 4096 1024 5; NXS_STFT_SPLIT_BYTES=1e12 timeout 200 python tools/run_stft.py $c 60 4096 1024 5; done
  timeout 200 python tools/run_stft.py 64 600 1024 256 5; NXS_STFT_SPLIT_BYTES=1e12 timeout 200 python tools/run_stft.py 64 600 1024 256 5; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
grep -q failed $OUT/pytest_stft.log && exit 1
timeout 2400 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 200 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
