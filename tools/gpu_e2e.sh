OUT=gpurun_out/r01c; mkdir -p $OUT
./tools/probe/pcie_probe 2>&1 | tee $OUT/pcie_probe2.txt
timeout 900 python -m pytest tests/test_stft_gpu.py -x -q > $OUT/pytest_stft.log 2>&1; tail -5 $OUT/pytest_stft.log
for kb in 1024 4096 16384 65536; do NXS_HOST_SLAB_KB=$kb timeout 300 python tools/run_e2e.py 8 600 3; done > $OUT/e2e.txt 2>&1
for t in 12 8; do NXS_HOST_THREADS=$t timeout 300 python tools/run_e2e.py 8 600 3; done >> $OUT/e2e.txt 2>&1
NXS_HOST_NO_MIRROR=1 timeout 300 python tools/run_e2e.py 8 600 3 >> $OUT/e2e.txt 2>&1
cat $OUT/e2e.txt
