#!/bin/bash
# r02c: real-packed overlap-save FIR kernel: parity tests + timings vs the pair kernel (variant 3)
OUT=gpurun_out/r02c; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fir_conv_gpu.py tests/test_full_size_gpu.py tests/test_stft_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
{ for v in 3 0 4; do echo "NXS_FIR_VARIANT=$v (3 = pair kernel 768 thr, 0 = real-packed 768 thr, 4 = real-packed 512 thr)"; NXS_FIR_VARIANT=$v timeout 200 python tools/run_fir.py 64 600 2049 5; done
for k in 1025 4097 10001; do for v in 3 0; do echo "K=$k variant $v"; NXS_FIR_VARIANT=$v timeout 200 python tools/run_fir.py 64 600 $k 3; done; done
timeout 120 python tools/run_stft.py 8 600 2048 512 10; } > $OUT/timings.txt 2>&1
cat $OUT/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_ols_r2c -s 2 -c 1 -o $OUT/fir_r2c_full -f python tools/run_fir.py 64 60 2049 2 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
