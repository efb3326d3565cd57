OUT=gpurun_out/r01l; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
{ for v in 0 5; do echo "stft variant $v"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 1024 256 20; done
for v in 0 1; do echo "istft variant $v"; NXS_ISTFT_VARIANT=$v timeout 120 python tools/run_istft.py 32 60 1024 256 20; done
timeout 200 python tools/run_mel.py 8 600 1024 256 128 0; } > $OUT/shapes.txt 2>&1
cat $OUT/shapes.txt
