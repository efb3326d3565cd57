# One gpurun call: c2r ISTFT parity tests, timings over shapes, ncu --set full of the 1024/256 kernel.
TAG=${1:-r01v}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_istft_c2r_gpu.py -x -q > $OUT/pytest_c2r.log 2>&1; tail -25 $OUT/pytest_c2r.log
{ for a in "32 60 1024 256" "32 60 1024 512" "32 60 1024 128" "32 60 512 128" "32 60 2048 512" "32 60 4096 1024"; do timeout 120 python tools/run_istft_c2r.py $a 10; done
timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10 514
timeout 120 python tools/run_istft.py 32 60 1024 256 10; } > $OUT/c2r_timings.txt 2>&1
cat $OUT/c2r_timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola_c2r -s 2 -c 1 -o $OUT/istft_c2r_full -f python tools/run_istft_c2r.py 32 60 1024 256 2 > $OUT/ncu_c2r.log 2>&1
tail -3 $OUT/ncu_c2r.log
