# One gpurun call: STFT + mel parity tests, STFT shape timings, log-mel timings, ncu of the fused kernel.
TAG=${1:-r01y}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_stft_gpu.py tests/test_stft_variants_gpu.py tests/test_mel_gpu.py tests/test_golden_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
{ for a in "8 600 1024 256" "128 60 4096 1024" "8 600 2048 512" "8 600 512 128" "32 60 8192 2048" "8 600 1024 250"; do timeout 120 python tools/run_stft.py $a 10; done
timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 48000; timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 16000; timeout 200 python tools/run_mel.py 8 600 1024 256 80 1 16000; } > $OUT/timings.txt 2>&1
cat $OUT/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 1 -c 1 -o $OUT/stft_mel_full -f python tools/run_mel.py 8 600 1024 256 128 0 48000 > $OUT/ncu_stft_mel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_r2c_staged -s 2 -c 1 -o $OUT/stft_full -f python tools/run_stft.py 8 600 1024 256 2 > $OUT/ncu_stft.log 2>&1
