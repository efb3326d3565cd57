OUT=gpurun_out/r01n; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_mel_gpu.py tests/test_golden_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
{ timeout 200 python tools/run_mel.py 8 600 1024 256 128 0; timeout 200 python tools/run_mel.py 32 60 1024 256 128 1;  timeout 200 python tools/run_mel.py 128 60 4096 1024 128 1; } > $OUT/mel.txt 2>&1
cat $OUT/mel.txt
