OUT=gpurun_out/r01g; mkdir -p $OUT
timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_stft_variants_gpu.py tests/test_istft_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
{ timeout 120 python tools/run_stft.py 128 60 4096 1024 10
timeout 120 python tools/run_stft.py 8 600 2048 512 10
timeout 120 python tools/run_stft.py 8 600 1024 256 10
timeout 120 python tools/run_stft.py 32 60 8192 2048 10
timeout 120 python tools/run_istft.py 32 60 1024 256 10
timeout 120 python tools/run_istft.py 32 60 512 128 10
timeout 120 python tools/run_istft.py 32 60 2048 512 10
timeout 120 python tools/run_istft.py 32 60 4096 1024 5; } > $OUT/shapes.txt 2>&1
cat $OUT/shapes.txt
