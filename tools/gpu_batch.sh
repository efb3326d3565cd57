OUT=gpurun_out/r01r; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_istft_gpu.py tests/test_golden_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
{ for h in 256 250 192 441 1024 64; do timeout 120 python tools/run_istft.py 32 60 1024 $h 10; done; timeout 120 python tools/run_istft.py 32 60 2048 500 5; timeout 120 python tools/run_istft.py 32 60 512 100 5; } > $OUT/odd_shapes.txt 2>&1
cat $OUT/odd_shapes.txt
