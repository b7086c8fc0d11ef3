#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_probe.py > gpurun_out/gpu_probe.txt 2> gpurun_out/gpu_probe.err; echo "probe rc=$?"; head -n 40 gpurun_out/gpu_probe.txt; tail -n 3 gpurun_out/gpu_probe.err
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "cosine or group_query or model or sqdiff or weights or host_api or ball_query_large" > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_r2b.log | tail -n 3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2b.log | head -n 30
python tools/model_profile.py > gpurun_out/model_profile.txt 2> gpurun_out/model_profile.err; echo "model_profile rc=$?"; tail -n 50 gpurun_out/model_profile.txt
cat gpurun_out/model_end_to_end_diff.json | head -c 1500
du -sh gpurun_out
