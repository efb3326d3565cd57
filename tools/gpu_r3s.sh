#!/bin/bash
# r03s: small transforms on the TMA-staged per-group kernel: nfft 256 (default now), 128 and 64 (variant 16): parity + timings
OUT=gpurun_out/r03s; mkdir -p $OUT
timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_stft_variants_gpu.py tests/test_golden_gpu.py tests/test_mel_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; echo "defaults: $(tail -1 $OUT/pytest.log)"
NXS_STFT_VARIANT=16 timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_golden_gpu.py -m gpu -q > $OUT/pytest_v16.log 2>&1; echo "variant 16: $(tail -1 $OUT/pytest_v16.log)"
{ for v in 0 19 18; do echo "nfft 256, NXS_STFT_VARIANT=$v (0 = staged 128 thr default, 19 = staged 256 thr, 18 = general kernel)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 256 64 10; done
  for v in 0 16; do echo "nfft 128 / 64, NXS_STFT_VARIANT=$v (0 = general kernel, 16 = staged)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 128 32 10; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 64 16 10; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 128 64 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
