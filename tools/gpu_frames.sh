TAG=${1:-r01S}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_frames_gpu.py tests/test_fir_conv_gpu.py tests/test_mel_gpu.py tests/test_golden_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
{ timeout 200 python tools/run_frames.py 5; timeout 200 python tools/run_fir.py 64 600 2049 5; timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 48000; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
