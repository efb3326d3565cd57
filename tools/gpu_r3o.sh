#!/bin/bash
# r03o: ring overlap-add ISTFT (any hop) and the nfft 512 STFT with the FFT engine on packed fp32x2: parity + timings against the scalar plans
OUT=gpurun_out/r03o; mkdir -p $OUT
NXS_ISTFT_RING_PK=1 timeout 900 python -m pytest tests/test_istft_gpu.py -m gpu -q > $OUT/pytest_ring_pk.log 2>&1; echo "ring PK: $(tail -1 $OUT/pytest_ring_pk.log)"
NXS_STFT_VARIANT=14 timeout 900 python -m pytest tests/test_stft_gpu.py -m gpu -q -k 512 > $OUT/pytest_stft512.log 2>&1; echo "stft 512 v14: $(tail -1 $OUT/pytest_stft512.log)"
{ for shape in "32 60 1024 250" "32 60 1024 441" "32 60 2048 700" "64 60 512 160" "64 60 256 100"; do echo "ISTFT ring $shape: scalar, packed"; timeout 120 python tools/run_istft.py $shape 10; NXS_ISTFT_RING_PK=1 timeout 120 python tools/run_istft.py $shape 10; done
  echo "STFT nfft 512: scalar, packed engine (variant 14)"; timeout 120 python tools/run_stft.py 8 600 512 128 10; NXS_STFT_VARIANT=14 timeout 120 python tools/run_stft.py 8 600 512 128 10; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
