#!/bin/bash
# r02i: select-based paired split pass (no divergent special case): parity + timings of the 4096 / 2048 variants
OUT=gpurun_out/r02i; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_stft_gpu.py tests/test_mel_gpu.py tests/test_stft_variants_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for v in 3 5 6; do NXS_STFT_VARIANT=$v timeout 600 python -m pytest tests/test_stft_gpu.py -m gpu -q > $OUT/pytest_v$v.log 2>&1; echo "variant $v: $(tail -1 $OUT/pytest_v$v.log)"; done
{ for v in 0 5 6 7 4; do echo "NXS_STFT_VARIANT=$v (0 = paired 256x2, 5 = paired XD 512x1, 6 = P32 384 thr, 7 = P32 256 thr, 4 = unpaired)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 128 60 4096 1024 10; done
for v in 0 3; do echo "NXS_STFT_VARIANT=$v (0 = unpaired 2048, 3 = paired)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 2048 512 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft4096_full -f python tools/run_stft.py 128 60 4096 1024 2 > $OUT/ncu.log 2>&1; tail -1 $OUT/ncu.log
