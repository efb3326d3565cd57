#!/bin/bash
# r03r: nfft 256 on the TMA-staged per-group kernel with half-warp groups (variants 16 / 17) against the general kernel
OUT=gpurun_out/r03r; mkdir -p $OUT
for v in 16 17; do NXS_STFT_VARIANT=$v timeout 600 python -m pytest tests/test_stft_gpu.py tests/test_stft_variants_gpu.py -m gpu -q -k "256 or hop or padding or cfg1" > $OUT/pytest_v$v.log 2>&1; echo "stft variant $v: $(tail -1 $OUT/pytest_v$v.log)"; done
{ for v in 0 16 17; do echo "NXS_STFT_VARIANT=$v"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 256 64 10; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 256 128 10; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 256 100 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
