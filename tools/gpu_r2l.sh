#!/bin/bash
# r02l: real-packed FIR kernel at N = 1024 for mid-size filters: parity + timings vs the pair kernel (variant 3)
OUT=gpurun_out/r02l; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fir_conv_gpu.py tests/test_host_pipeline_gpu.py tests/test_c_abi.py -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
{ for k in 131 255 385 513; do for v in 3 0; do echo "K=$k variant $v (3 = pair kernel, 0 = real-packed)"; NXS_FIR_VARIANT=$v timeout 200 python tools/run_fir.py 64 600 $k 3; done; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
