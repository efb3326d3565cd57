#!/bin/bash
# r02a: paired split pass (nfft 2048 / 4096) -- parity tests, old-vs-new timings, box probe
OUT=gpurun_out/r02a; mkdir -p $OUT
{ nproc; lscpu | grep -E "Model name|Socket|NUMA|CPU\(s\)"; free -g | head -2; nvidia-smi topo -m; cat /sys/devices/system/node/online; } > $OUT/box.txt 2>&1
timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_mel_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
{ for v in 4 0; do echo "NXS_STFT_VARIANT=$v (4 = unpaired 4096)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 128 60 4096 1024 10; done
for v in 3 0; do echo "NXS_STFT_VARIANT=$v (3 = unpaired 2048)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 2048 512 10; done
timeout 120 python tools/run_stft.py 8 600 1024 256 10
timeout 120 python tools/run_stft.py 32 60 8192 2048 10; } > $OUT/timings.txt 2>&1
cat $OUT/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft4096_full -f python tools/run_stft.py 128 60 4096 1024 2 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
