TAG=${1:-r01B}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_postops_gpu.py -x -q > $OUT/pytest_post.log 2>&1; tail -5 $OUT/pytest_post.log
timeout 600 python tools/run_post.py > $OUT/post_timings.txt 2>&1; cat $OUT/post_timings.txt
