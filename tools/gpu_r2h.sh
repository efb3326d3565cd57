#!/bin/bash
# r02h: full GPU suite, XD (double exchange buffer, 512 threads) vs the paired 4096 STFT, compute-sanitizer
OUT=gpurun_out/r02h; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
{ for v in 0 5 4; do echo "NXS_STFT_VARIANT=$v (0 = paired 256x2, 5 = paired XD 512x1, 4 = unpaired)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 128 60 4096 1024 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
NXS_STFT_VARIANT=5 timeout 600 python -m pytest tests/test_stft_gpu.py -m gpu -q -k "4096 or plans or pow2" > $OUT/pytest_xd.log 2>&1; tail -3 $OUT/pytest_xd.log
bash tools/gpu_sanitize.sh > $OUT/sanitize.log 2>&1; tail -8 $OUT/sanitize.log; cp gpurun_out/sanitize/*.log $OUT/ 2>/dev/null
