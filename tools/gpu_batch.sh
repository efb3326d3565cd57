OUT=gpurun_out/r01s; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_fir_conv_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
{ for k in 2049 2731 2732 3585 4097 8193 16385; do timeout 200 python tools/run_fir.py 64 600 $k 2; done; } > $OUT/fir.txt 2>&1
cat $OUT/fir.txt
