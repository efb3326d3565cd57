OUT=gpurun_out/r01q; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_stft_gpu.py tests/test_stft_variants_gpu.py tests/test_mel_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
{ for h in 256 250 441 192; do timeout 120 python tools/run_stft.py 8 600 1024 $h 10; done; timeout 120 python tools/run_stft.py 128 60 4096 1001 5; } > $OUT/odd_shapes.txt 2>&1
cat $OUT/odd_shapes.txt
