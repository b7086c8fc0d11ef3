#!/bin/bash
# first-level GPU check of a round: GPU tests, smoke, a short bench line, the model profile
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?"
tail -n 30 gpurun_out/pytest_r2.log
python __graft_entry__.py --smoke > gpurun_out/smoke_r2.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke_r2.log
python tools/model_profile.py > gpurun_out/model_profile.txt 2> gpurun_out/model_profile.err; echo "model_profile rc=$?"; head -n 30 gpurun_out/model_profile.txt; tail -n 5 gpurun_out/model_profile.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_r2.err; head -c 6000 gpurun_out/bench_r2.json
