#!/bin/bash
# r03b: ISTFT (c64 and c2r) register-overlap-add plans on packed fp32x2 (default) against NXS_ISTFT_SCALAR=1, per shape; parity of the packed plans
OUT=gpurun_out/r03b; mkdir -p $OUT
timeout 900 python -m pytest tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py tests/test_stft_gpu.py tests/test_full_size_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; echo "packed defaults: $(tail -1 $OUT/pytest.log)"
{
  for shape in "64 60 256 64" "64 60 512 128" "32 60 1024 512" "32 60 1024 128" "32 60 1024 256" "32 60 2048 512" "32 60 2048 1024" "16 60 4096 1024"; do
    echo "ISTFT shape $shape: packed, scalar"; timeout 120 python tools/run_istft.py $shape 10; NXS_ISTFT_SCALAR=1 timeout 120 python tools/run_istft.py $shape 10
  done
  for shape in "64 60 512 128" "32 60 1024 256" "32 60 1024 512" "32 60 2048 512" "16 60 4096 1024"; do
    echo "c2r shape $shape: packed, scalar"; timeout 120 python tools/run_istft_c2r.py $shape 10; NXS_ISTFT_SCALAR=1 timeout 120 python tools/run_istft_c2r.py $shape 10
  done
  echo "STFT 2048 packed engine (default) / scalar (variant 9)"; timeout 120 python tools/run_stft.py 64 60 2048 512 10; NXS_STFT_VARIANT=9 timeout 120 python tools/run_stft.py 64 60 2048 512 10
} > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
