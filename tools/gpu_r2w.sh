#!/bin/bash
# r02w: warp-per-frame ISTFT as the default, small-call single-stream host path: full GPU suite, smoke, bench + reference arm
OUT=gpurun_out/r02w; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02w/bench_n1.json"))
print(json.dumps(d["cfg1"], indent=1))
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"])
print(json.dumps(d["other_kernels"], indent=1)[:3000])
PY
