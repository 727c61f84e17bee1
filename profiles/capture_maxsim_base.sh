#!/bin/bash
# gpurun -- 'bash profiles/capture_maxsim_base.sh <tag>': the MaxSim single-query kernel under `--clock-control base` (SM
# clocks locked low, like the power-capped sustained regime) with the hi pass's A operand in tensor memory
# (INNR_MAXSIM_TS=1) and in shared memory (=0): raw metrics + per-instruction stall samples as CSV.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for ts in 1 0; do
  INNR_MAXSIM_TS=$ts ncu --set full --clock-control base --import-source on -k regex:maxsim_tc_kernel -s 2 -c 1 -f \
      -o $OUT/${TAG}_ts${ts} python innr_b200/csrc/dev/maxsim_prof.py > $OUT/${TAG}_ts${ts}_ncu.log 2>&1
  ncu -i $OUT/${TAG}_ts${ts}.ncu-rep --page raw --csv > $OUT/${TAG}_ts${ts}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_ts${ts}.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_ts${ts}_source.csv 2>/dev/null
  python profiles/top_sass.py < $OUT/${TAG}_ts${ts}_source.csv > $OUT/${TAG}_ts${ts}_top_sass.txt 2>&1
  gzip -f $OUT/${TAG}_ts${ts}_source.csv
  rm -f $OUT/${TAG}_ts${ts}.ncu-rep
done
ls -la $OUT | grep ${TAG}_ts
