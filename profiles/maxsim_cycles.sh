#!/bin/bash
# gpurun -- 'bash profiles/maxsim_cycles.sh': SM cycles of the single-query MaxSim launch at locked base clocks for the two
# operand placements and the timing-only debug variants (bit 8: no Xhi tcgen05.st, 16: no Xlo st, 32: no converter LDS,
# 64: no epilogue tcgen05.ld, 4: no epilogue math). Results of the debug variants are wrong by construction.
for ts in 1 0; do
  for dbg in 0 8 16 24 32 64 4; do
    if [ $ts = 0 ] && [ $dbg = 8 -o $dbg = 24 ]; then continue; fi
    INNR_MAXSIM_TS=$ts INNR_MAXSIM_DEBUG=$dbg ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control base \
        -k regex:maxsim_tc_kernel -s 2 -c 1 --csv python innr_b200/csrc/dev/maxsim_prof.py 2>/dev/null | grep maxsim_tc | \
        awk -F'","' -v ts=$ts -v dbg=$dbg '{gsub(/"/,"",$NF); printf "ts=%s dbg=%s %s %s\n", ts, dbg, $(NF-2), $NF}'
  done
done
