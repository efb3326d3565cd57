#!/bin/bash
# r02v: warp-per-frame ISTFT with the overlap-add carry in shared memory: parity + timings; octet-wise twiddle application everywhere: parity
OUT=gpurun_out/r02v; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_stft_gpu.py tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py tests/test_fir_conv_gpu.py tests/test_mel_gpu.py tests/test_golden_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
for v in 5 6 7 8; do NXS_ISTFT_VARIANT=$v timeout 600 python -m pytest tests/test_istft_gpu.py tests/test_full_size_gpu.py -m gpu -q -k "istft or cfg5" > $OUT/pytest_v$v.log 2>&1; echo "variant $v: $(tail -1 $OUT/pytest_v$v.log)"; done
{ for v in 0 3 5 6 7 8; do echo "NXS_ISTFT_VARIANT=$v (0 = default T=64 XD, 3 = warp/frame regs 256 thr, 5/6/7/8 = warp/frame smem carry 384/448/320/352 thr)"; NXS_ISTFT_VARIANT=$v timeout 120 python tools/run_istft.py 32 60 1024 256 10; done
timeout 120 python tools/run_stft.py 8 600 1024 256 10; timeout 120 python tools/run_stft.py 128 60 4096 1024 10; timeout 200 python tools/run_fir.py 64 600 2049 3; timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
