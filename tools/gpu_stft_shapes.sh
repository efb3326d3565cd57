OUT=gpurun_out/r01f; mkdir -p $OUT
timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_stft_variants_gpu.py -x -q > $OUT/pytest_stft.log 2>&1; tail -5 $OUT/pytest_stft.log
for v in 0 2 3; do echo "variant $v"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 128 60 4096 1024 10; done > $OUT/stft_shapes.txt 2>&1
for v in 0 2; do echo "variant $v"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 2048 512 10; done >> $OUT/stft_shapes.txt 2>&1
for v in 0 4; do echo "variant $v"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 1024 256 10; done >> $OUT/stft_shapes.txt 2>&1
for v in 0 1; do echo "variant $v"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 32 60 8192 2048 10; done >> $OUT/stft_shapes.txt 2>&1
cat $OUT/stft_shapes.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft4096_full -f python tools/run_stft.py 128 60 4096 1024 2 > $OUT/ncu_stft.log 2>&1
