OUT=gpurun_out/r01h; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fir_conv_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
{ for v in 0 1; do echo "variant $v"; NXS_FIR_VARIANT=$v timeout 200 python tools/run_fir.py 64 600 2049 5; NXS_FIR_VARIANT=$v timeout 200 python tools/run_fir.py 64 600 255 5; done
echo "old kernel"; NXS_FIR_NO_PG=1 timeout 200 python tools/run_fir.py 64 600 2049 5; } > $OUT/fir.txt 2>&1
cat $OUT/fir.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_ols_pg -s 2 -c 1 -o $OUT/fir_pg_full -f python tools/run_fir.py 64 60 2049 2 > $OUT/ncu_fir.log 2>&1
