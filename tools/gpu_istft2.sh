# One gpurun call: ISTFT (c64 + c2r) parity tests, timings, ncu --set full of both register overlap-add kernels.
TAG=${1:-r01z}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py tests/test_golden_gpu.py -x -q > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
{ for a in "32 60 1024 256" "32 60 512 128" "32 60 2048 512" "32 60 4096 1024" "32 60 1024 512" "32 60 1024 128" "32 60 1024 250"; do timeout 120 python tools/run_istft.py $a 10; done
for a in "32 60 1024 256" "32 60 1024 512" "32 60 1024 128" "32 60 512 128" "32 60 2048 512" "32 60 4096 1024"; do timeout 120 python tools/run_istft_c2r.py $a 10; done; } > $OUT/timings.txt 2>&1
cat $OUT/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola_kernel -s 2 -c 1 -o $OUT/istft_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu_istft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola_c2r -s 2 -c 1 -o $OUT/istft_c2r_full -f python tools/run_istft_c2r.py 32 60 1024 256 2 > $OUT/ncu_c2r.log 2>&1
