#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "cosine or group or interpolate or emd or index_points or model" > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_r2d.log | tail -n 3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2d.log | head -n 40
grep -E "^E  " gpurun_out/pytest_r2d.log | head -n 20
python bench.py --steps 20 --warmup 5 --no-model --no-eval > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r2d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
for k,v in d['kernels'].items():
    print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note','levels')})
PY
