#!/bin/bash
# r03z: last check of the final tree: the STFT tests (incl. the channel-block walk) and smoke
OUT=gpurun_out/r03z; mkdir -p $OUT
timeout 100 python -m pytest tests/test_stft_gpu.py -m gpu -q -x > $OUT/pytest_stft.log 2>&1; echo "stft: $(tail -1 $OUT/pytest_stft.log)"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
