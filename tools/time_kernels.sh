#!/bin/bash
# Developer tool: per-kernel times (natural clocks) of one knn_point(k) step for several library builds.
#   tools/time_kernels.sh k lib1.so lib2.so ...
k=$1; shift
for v in "$@"; do
  echo "== $v"
  B200PCI_LIB=$v ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"knn_|nbr_" -s 6 -c 6 python tools/prof_one.py knn $k 8 2>&1 \
    | grep -E "^  [a-z_ ]*(knn|nbr)|gpu__time" | sed -E "s/\(NbrParams.*//; s/\(int.*//" | paste - - | awk '{printf "   %-40s %s %s\n", $1" "$2, $(NF-1), $NF}'
done
