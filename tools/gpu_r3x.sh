#!/bin/bash
# r03x: one large call against shard-sized calls over the same tensors (footprint dependence of the achieved rate)
OUT=gpurun_out/r03x; mkdir -p $OUT
{ for c in 256 512 1024; do timeout 200 python tools/run_split_calls.py $c; done; } > $OUT/split.txt 2>&1; cat $OUT/split.txt
