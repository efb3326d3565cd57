OUT=gpurun_out/r01k; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_mel_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
{ timeout 200 python tools/run_mel.py 8 600 1024 256 128 0; timeout 200 python tools/run_mel.py 128 60 4096 1024 128 1; } > $OUT/mel.txt 2>&1
cat $OUT/mel.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 4 -c 1 -o $OUT/stftmel_full -f python tools/run_mel.py 8 600 1024 256 128 0 > $OUT/ncu.log 2>&1
