OUT=gpurun_out/r01i; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
{ timeout 200 python tools/run_fir.py 64 600 2049 5; timeout 200 python tools/run_fir.py 64 600 255 5
timeout 120 python tools/run_istft.py 32 60 1024 256 10; } > $OUT/shapes.txt 2>&1
cat $OUT/shapes.txt
