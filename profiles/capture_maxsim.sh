#!/bin/bash
# gpurun -- 'bash profiles/capture_maxsim.sh <tag>': plain timings of the MaxSim single-query and query-batch kernels, then one
# ncu --set full capture of each (kernel ids 2 = third single-query launch, 3 = first batch launch), exported to CSV
# (raw metrics + per-instruction SASS stall samples) so the summaries can be read without the .ncu-rep.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python innr_b200/csrc/dev/maxsim_prof.py > $OUT/${TAG}_maxsim_plain.log 2>&1 || { tail -5 $OUT/${TAG}_maxsim_plain.log; exit 1; }
cat $OUT/${TAG}_maxsim_plain.log
ncu --set full --clock-control none --import-source on -k regex:maxsim_tc_kernel -s 2 -c 2 -f -o $OUT/${TAG}_maxsim_full \
    python innr_b200/csrc/dev/maxsim_prof.py > $OUT/${TAG}_maxsim_ncu.log 2>&1
ncu -i $OUT/${TAG}_maxsim_full.ncu-rep --page raw --csv > $OUT/${TAG}_maxsim_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_maxsim_full.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_maxsim_source.csv 2>/dev/null
ls -la $OUT | grep ${TAG}_maxsim
