TAG=${1:-r01B}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_postops_gpu.py -x -q > $OUT/pytest_post.log 2>&1; tail -40 $OUT/pytest_post.log
