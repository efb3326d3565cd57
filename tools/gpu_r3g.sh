#!/bin/bash
# r03g: FIR real-packed kernel with two warps per block and radices 64 x 64 (variant 7) against the default; c2r ISTFT with the packed pre-pass / overlap-add
OUT=gpurun_out/r03g; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fir_conv_gpu.py tests/test_istft_c2r_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
NXS_FIR_VARIANT=7 timeout 900 python -m pytest tests/test_fir_conv_gpu.py tests/test_full_size_gpu.py tests/test_host_pipeline_gpu.py -m gpu -q -k "fir or conv or cfg4" > $OUT/pytest_v7.log 2>&1; echo "FIR variant 7: $(tail -1 $OUT/pytest_v7.log)"
{ for v in 0 7 0 7; do echo "NXS_FIR_VARIANT=$v"; for k in 2049 513 4097; do NXS_FIR_VARIANT=$v timeout 200 python tools/run_fir.py 64 600 $k 3; done; done
  for a in "32 60 1024 256" "32 60 1024 512" "32 60 2048 512" "64 60 512 128"; do timeout 120 python tools/run_istft_c2r.py $a 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
