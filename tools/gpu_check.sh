#!/bin/bash
# first-level GPU check of a round: GPU tests, smoke, a short bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_r2.log | tail -n 3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2.log | head -n 40
python __graft_entry__.py --smoke > gpurun_out/smoke_r2.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_r2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_r2.err; head -c 3000 gpurun_out/bench_r2.json
python -c "
import json
d=json.load(open('gpurun_out/model_shadow_stats.json')); print(d['calls']); print(d['failures'][:8])
print(open('gpurun_out/model_end_to_end_diff.json').read()[-700:])"
du -sh gpurun_out
