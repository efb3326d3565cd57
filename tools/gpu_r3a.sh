#!/bin/bash
# (historical: NXS_ISTFT_PK was replaced by packed defaults with NXS_ISTFT_SCALAR=1 as the A/B switch)
# r03a: FFT engine on packed fp32x2 (Plan::PK) for the remaining transform kernels: parity + timings per variant
OUT=gpurun_out/r03a; mkdir -p $OUT
for v in 14 15; do NXS_STFT_VARIANT=$v timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_mel_gpu.py tests/test_golden_gpu.py -m gpu -q > $OUT/pytest_stft_v$v.log 2>&1; echo "stft variant $v: $(tail -1 $OUT/pytest_stft_v$v.log)"; done
NXS_ISTFT_PK=1 timeout 900 python -m pytest tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py -m gpu -q > $OUT/pytest_istft_pk.log 2>&1; echo "istft PK: $(tail -1 $OUT/pytest_istft_pk.log)"
{
  for v in 0 14 15; do echo "NXS_STFT_VARIANT=$v (0 default scalar, 14 engine packed, 15 engine + window packed)"
    NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 1024 256 10
    NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 64 60 2048 512 10
    NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 32 60 8192 2048 10
    NXS_STFT_VARIANT=$v timeout 120 python tools/run_mel.py 2>/dev/null | tail -1
  done
  timeout 120 python tools/run_stft.py 128 60 4096 1024 10
  echo "ISTFT 2048 / c2r 1024: default, then NXS_ISTFT_PK=1"
  timeout 120 python tools/run_istft.py 32 60 2048 512 10; timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10
  NXS_ISTFT_PK=1 timeout 120 python tools/run_istft.py 32 60 2048 512 10; NXS_ISTFT_PK=1 timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10
  timeout 120 python tools/run_istft.py 32 60 1024 256 10
  timeout 200 python tools/run_fir.py 64 600 2049 3
} > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
