#!/bin/bash
# r02r: mixed transfer mode of the host STFT entry: parity of every mode / memory kind, e2e timings per mode
OUT=gpurun_out/r02r; mkdir -p $OUT
timeout 900 python -m pytest tests/test_host_pipeline_gpu.py tests/test_stft_gpu.py tests/test_c_abi.py -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-multi --no-extras --no-cpu --e2e-steps 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02r/bench_n1.json")); e = d["e2e"]
print({k: e[k] for k in e if k.startswith("ms_per_step") or k in ("transfer_mode_chosen", "host_timeline_ms")})
PY
