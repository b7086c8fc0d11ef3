#!/bin/bash
# ncu evidence; reports are exported to CSV on the box and deleted (gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
SECTIONS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats"
python tools/prof_ops.py knn16 > gpurun_out/prof_knn16_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "prof/" -c 12 -f -o /tmp/r2_knn16 python tools/prof_ops.py knn16 > gpurun_out/ncu_knn16.log 2>&1
echo "ncu knn16 rc=$?"; tail -n 2 gpurun_out/ncu_knn16.log
ncu -i /tmp/r2_knn16.ncu-rep --page raw --csv > gpurun_out/r2_knn16_raw.csv 2>/dev/null
ncu -i /tmp/r2_knn16.ncu-rep --page source --csv --kernel-name regex:knn_scan_tc > gpurun_out/r2_knn_scan_tc_source.csv 2>/dev/null
ls -la /tmp/r2_knn16.ncu-rep gpurun_out/r2_knn16_raw.csv gpurun_out/r2_knn_scan_tc_source.csv
python tools/prof_ops.py > gpurun_out/prof_ops_plain.log 2>&1 && \
ncu $SECTIONS --clock-control none --nvtx --nvtx-include "prof/" -c 110 -f -o /tmp/r2_ops python tools/prof_ops.py knn32_b1 chamfer group128 group3 rows_gather gather interp three_nn ball fps fps_b1 knn_mid knn3_mid sqdiff emd4096 cosine group_concat > gpurun_out/ncu_ops.log 2>&1
echo "ncu ops rc=$?"; tail -n 2 gpurun_out/ncu_ops.log
ncu -i /tmp/r2_ops.ncu-rep --page raw --csv > gpurun_out/r2_ops_raw.csv 2>/dev/null
ls -la /tmp/r2_ops.ncu-rep gpurun_out/r2_ops_raw.csv
python bench.py --steps 2 --warmup 3 --no-extras --no-model --no-eval --no-parity > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_knn16_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-model --no-eval --no-parity > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
du -sh gpurun_out
