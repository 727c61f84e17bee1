#!/bin/bash
# Run on the GPU box (gpurun -- 'bash profiles/capture.sh <tag>'): per workload, the plain bench run, then the ncu launch
# list (gpu__time_duration.sum, cold-cache + serialised: compare SHARES) and one `ncu --set full` capture of the dominant
# kernel. Outputs land in gpurun_out/<tag>_*; the summaries worth judging are copied into profiles/ by hand
# (profiles/summarise.py turns the .ncu-rep files into the CSV rows committed there).
set -u
TAG=${1:-r01}
WORKLOADS=${2:-"knn_cosine_1q hamming u8 maxsim knn_cosine_multi batch_demo"}
OUT=gpurun_out
mkdir -p $OUT
declare -A KERN=( [knn_cosine_1q]=pdx_scan_kernel [hamming]=hamming_kernel [u8]=u8_scan_kernel [maxsim]=maxsim_tc_kernel
                  [knn_cosine_multi]=knn_tc_filter_kernel [batch_demo]=pdx_scan )
declare -A COUNT=( [knn_cosine_1q]=2 [hamming]=2 [u8]=2 [maxsim]=2 [knn_cosine_multi]=8 [batch_demo]=2 )
# the driver's own command (all workloads in one process): plain run, then its ncu launch list -- the kernels' SHARES of a
# step must agree with bench.py's live CUDA-event numbers (per-launch times under ncu are cold-cache and serialised)
DEF="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$DEF > $OUT/${TAG}_plain_default.log 2>&1 || { echo "plain default run failed"; tail -5 $OUT/${TAG}_plain_default.log; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_default.csv $DEF \
    > $OUT/${TAG}_launches_default.log 2>&1
echo "default command: done"
for w in $WORKLOADS; do
  CMD="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline"
  $CMD > $OUT/${TAG}_plain_$w.log 2>&1 || { echo "plain run of $w failed"; tail -5 $OUT/${TAG}_plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_$w.csv $CMD \
      > $OUT/${TAG}_launches_$w.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:${KERN[$w]} -s 3 -c ${COUNT[$w]} -f -o $OUT/${TAG}_full_$w $CMD \
      > $OUT/${TAG}_full_$w.log 2>&1
  # gpurun brings back at most 64 MiB: export what the summaries need (all raw metrics; per-instruction samples and
  # shared-memory wavefronts) and drop the report itself
  ncu -i $OUT/${TAG}_full_$w.ncu-rep --page raw --csv > $OUT/${TAG}_full_$w.raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_full_$w.ncu-rep --page source --csv --print-source sass 2>/dev/null | \
      python profiles/top_sass.py > $OUT/${TAG}_full_$w.top_sass.txt
  rm -f $OUT/${TAG}_full_$w.ncu-rep
  echo "$w: done"
done
ls -la $OUT | tail -30
