#!/bin/bash
# r02s: e2e of the host STFT entry vs the number of host threads (same box, back to back)
OUT=gpurun_out/r02s; mkdir -p $OUT
{ for t in 16 12 10 8 6 4; do echo "NXS_HOST_THREADS=$t"; NXS_HOST_THREADS=$t timeout 200 python tools/run_e2e.py 8 600 4; done; } > $OUT/e2e_threads.txt 2>&1; cat $OUT/e2e_threads.txt
