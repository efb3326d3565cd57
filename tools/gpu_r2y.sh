#!/bin/bash
# r02y: packed fp32x2 arithmetic per plan (Plan::PK) in the FIR, warp-per-frame ISTFT and nfft 4096 STFT kernels: parity + timings
OUT=gpurun_out/r02y; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_stft_gpu.py tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py tests/test_fir_conv_gpu.py tests/test_mel_gpu.py tests/test_golden_gpu.py tests/test_host_pipeline_gpu.py tests/test_full_size_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
{
  timeout 120 python tools/run_stft.py 8 600 1024 256 10
  echo "nfft 4096 packed (default) / scalar (NXS_STFT_VARIANT=9)"
  timeout 120 python tools/run_stft.py 128 60 4096 1024 10; NXS_STFT_VARIANT=9 timeout 120 python tools/run_stft.py 128 60 4096 1024 10
  timeout 120 python tools/run_stft.py 128 60 4096 1024 10; NXS_STFT_VARIANT=9 timeout 120 python tools/run_stft.py 128 60 4096 1024 10
  echo "ISTFT packed warp-per-frame (default) / scalar T=64 two buffers (NXS_ISTFT_VARIANT=2)"
  timeout 120 python tools/run_istft.py 32 60 1024 256 10; NXS_ISTFT_VARIANT=2 timeout 120 python tools/run_istft.py 32 60 1024 256 10
  timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10
  echo "FIR packed"
  timeout 200 python tools/run_fir.py 64 600 2049 3; timeout 200 python tools/run_fir.py 64 600 255 3; timeout 200 python tools/run_fir.py 64 600 513 3; timeout 200 python tools/run_fir.py 64 600 8191 3
} > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
