#!/bin/bash
# r02o: as_windowed (4 loads in flight, streaming stores), FIR interior element stores; parity + timings
OUT=gpurun_out/r02o; mkdir -p $OUT
timeout 900 python -m pytest tests/test_frames_gpu.py tests/test_fir_conv_gpu.py tests/test_golden_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
{ timeout 200 python tools/run_frames.py 10; for k in 255 257 385 2048 2049; do echo "K=$k"; timeout 200 python tools/run_fir.py 64 600 $k 3; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
