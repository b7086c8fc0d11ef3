#!/bin/bash
# ncu evidence; reports are exported to CSV on the box and deleted (gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
SECTIONS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats"
OPS="${PROF_OPS:-group128 group3 rows_gather gather interp cosine group_concat emd4096 three_nn}"
python tools/prof_ops.py $OPS > gpurun_out/prof_ops_plain.log 2>&1 && \
ncu $SECTIONS --clock-control none --nvtx --nvtx-include "prof/" -c 110 -f -o /tmp/r2_ops python tools/prof_ops.py $OPS > gpurun_out/ncu_ops.log 2>&1
echo "ncu ops rc=$?"; tail -n 2 gpurun_out/ncu_ops.log
ncu -i /tmp/r2_ops.ncu-rep --page raw --csv > gpurun_out/r2_ops2_raw.csv 2>/dev/null; ls -la gpurun_out/r2_ops2_raw.csv
# the bench's own launch of the dominant kernel (64 pairs): DRAM traffic for roofline.traffic
python bench.py --steps 1 --warmup 3 --no-extras --no-model --no-eval --no-parity > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --set full --clock-control none -k regex:knn_scan_tc --launch-skip 2 -c 1 -f -o /tmp/r2_scan64 python bench.py --steps 1 --warmup 3 --no-extras --no-model --no-eval --no-parity > gpurun_out/ncu_scan64.log 2>&1
echo "ncu scan64 rc=$?"; tail -n 2 gpurun_out/ncu_scan64.log
ncu -i /tmp/r2_scan64.ncu-rep --page raw --csv > gpurun_out/r2_scan_B64_raw.csv 2>/dev/null; ls -la gpurun_out/r2_scan_B64_raw.csv
du -sh gpurun_out
