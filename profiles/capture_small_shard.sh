#!/bin/bash
# gpurun -- 'bash profiles/capture_small_shard.sh <tag>': the C2a and C4 scans on a 1/8 shard (what one of 8 GPUs holds), one
# ncu --set full capture each, raw metrics exported to CSV: where does the 4 % go that a 1/8 shard loses against 1/8 of the
# full-size kernel time?
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for w in knn_cosine_1q hamming; do
  K=pdx_scan_kernel; [ $w = hamming ] && K=hamming_kernel
  CMD="python bench.py --workload $w --scale 0.125 --steps 3 --warmup 3 --no-cpu-baseline"
  ncu --set full --clock-control none -k regex:$K -s 3 -c 2 -f -o $OUT/${TAG}_small_$w $CMD > $OUT/${TAG}_small_$w.log 2>&1
  ncu -i $OUT/${TAG}_small_$w.ncu-rep --page raw --csv > $OUT/${TAG}_small_$w.raw.csv 2>/dev/null
  rm -f $OUT/${TAG}_small_$w.ncu-rep
done
ls -la $OUT | grep ${TAG}_small
