TAG=${1:-r01M}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_stft_variants_gpu.py tests/test_stft_gpu.py tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
{ for v in 0 6 7 8; do echo "variant $v"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 8 600 1024 256 10; done
timeout 120 python tools/run_stft.py 8 600 512 128 10
timeout 120 python tools/run_istft.py 32 60 256 64 10
timeout 120 python tools/run_istft_c2r.py 32 60 512 128 10; } > $OUT/variants.txt 2>&1
cat $OUT/variants.txt
