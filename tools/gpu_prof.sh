#!/bin/bash
# probes + ncu evidence (one ncu "use" per call: all ncu runs here count as one)
mkdir -p gpurun_out
python tools/gpu_probe.py > gpurun_out/gpu_probe.txt 2> gpurun_out/gpu_probe.err; echo "probe rc=$?"; cat gpurun_out/gpu_probe.txt; tail -n 5 gpurun_out/gpu_probe.err
python tools/prof_ops.py > gpurun_out/prof_ops_plain.log 2>&1 && \
ncu --set full --clock-control none --nvtx --nvtx-include "prof/" -c 100 -f -o gpurun_out/r2_ops python tools/prof_ops.py > gpurun_out/ncu_ops.log 2>&1
echo "ncu ops rc=$?"; tail -n 3 gpurun_out/ncu_ops.log; ls -la gpurun_out/r2_ops.ncu-rep
python bench.py --steps 2 --warmup 3 --no-extras --no-model --no-eval --no-parity > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_knn16_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-model --no-eval --no-parity > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
