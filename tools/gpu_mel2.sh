# One gpurun call: mel parity tests, log-mel timings at two sampling rates, ncu --set full of the fused kernel.
TAG=${1:-r01w}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_mel_gpu.py -x -q > $OUT/pytest_mel.log 2>&1; tail -25 $OUT/pytest_mel.log
{ timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 48000; timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 16000; timeout 200 python tools/run_mel.py 8 600 1024 256 80 1 16000; } > $OUT/mel_timings.txt 2>&1
cat $OUT/mel_timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 1 -c 1 -o $OUT/stft_mel_full -f python tools/run_mel.py 8 600 1024 256 128 0 48000 > $OUT/ncu_stft_mel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 1 -c 1 -o $OUT/stft_mel16k_full -f python tools/run_mel.py 8 600 1024 256 128 0 16000 > $OUT/ncu_stft_mel16k.log 2>&1
