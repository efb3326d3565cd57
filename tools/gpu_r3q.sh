#!/bin/bash
# r03q: small transforms (nfft 64 / 128 / 256): where they stand, and the FFT engine on packed fp32x2 (variant 14)
OUT=gpurun_out/r03q; mkdir -p $OUT
NXS_STFT_VARIANT=14 timeout 600 python -m pytest tests/test_stft_gpu.py -m gpu -q > $OUT/pytest_v14.log 2>&1; echo "stft variant 14: $(tail -1 $OUT/pytest_v14.log)"
{ for shape in "8 600 256 64" "8 600 128 32" "8 600 64 16" "8 600 256 128"; do echo "STFT $shape: scalar, packed"; timeout 120 python tools/run_stft.py $shape 10; NXS_STFT_VARIANT=14 timeout 120 python tools/run_stft.py $shape 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
