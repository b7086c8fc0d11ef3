#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "bitwise or model or permuted or sqdiff or host_api or golden" > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_r2c.log | tail -n 3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2c.log | head -n 40
grep -E "^E  " gpurun_out/pytest_r2c.log | head -n 20
python -c "
import json
d=json.load(open('gpurun_out/model_shadow_stats.json')); print(d['calls']); print(d['failures'][:8])"
