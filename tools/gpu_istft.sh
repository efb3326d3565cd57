OUT=gpurun_out/r01d; mkdir -p $OUT
timeout 900 python -m pytest tests/test_istft_gpu.py -x -q > $OUT/pytest_istft.log 2>&1; tail -15 $OUT/pytest_istft.log
timeout 120 python tools/run_istft.py 32 60 1024 256 10 > $OUT/istft.txt 2>&1
NXS_ISTFT_NO_ROLA=1 timeout 120 python tools/run_istft.py 32 60 1024 256 10 >> $OUT/istft.txt 2>&1
timeout 120 python tools/run_istft.py 32 60 1024 512 10 >> $OUT/istft.txt 2>&1
timeout 120 python tools/run_istft.py 32 60 512 128 10 >> $OUT/istft.txt 2>&1
timeout 120 python tools/run_istft.py 32 60 2048 512 10 >> $OUT/istft.txt 2>&1
timeout 120 python tools/run_istft.py 32 60 4096 1024 5 >> $OUT/istft.txt 2>&1
cat $OUT/istft.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola -s 2 -c 1 -o $OUT/istft_rola_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu_istft.log 2>&1
