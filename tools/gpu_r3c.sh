#!/bin/bash
# r03c: FIR real-packed kernel with the index-reversal inverse + slim stores: parity, timings; then validation of the whole tree
OUT=gpurun_out/r03c; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fir_conv_gpu.py tests/test_host_pipeline_gpu.py tests/test_full_size_gpu.py -m gpu -q -k "fir or conv or cfg4" > $OUT/pytest_fir.log 2>&1; tail -2 $OUT/pytest_fir.log
{ timeout 200 python tools/run_fir.py 64 600 2049 3; timeout 200 python tools/run_fir.py 64 600 513 3; timeout 200 python tools/run_fir.py 64 600 385 3; timeout 200 python tools/run_fir.py 64 600 255 3; timeout 200 python tools/run_fir.py 64 600 8191 3; } > $OUT/fir_timings.txt 2>&1; cat $OUT/fir_timings.txt
grep -q "failed" $OUT/pytest_fir.log && exit 1
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03c/bench_n1.json"))
print(json.dumps(d["cfg1"], indent=1))
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"])
for k, v in d["other_kernels"].items():
    if isinstance(v, dict): print(k, {kk: vv for kk, vv in v.items() if kk in ("kernel_ms", "frac_of_hbm_peak", "ms_per_step")})
print(json.dumps(d.get("multi_gpu"), indent=1)[:1500])
PY
