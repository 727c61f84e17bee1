for cfg in "0 1" "0 2" "1 1"; do set -- $cfg
  INNR_MAXSIM_TS=$1 INNR_MAXSIM_DEBUG=$2 ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,dram__bytes_read.sum --clock-control base \
        -k regex:maxsim_tc_kernel -s 2 -c 1 --csv python innr_b200/csrc/dev/maxsim_prof.py 2>/dev/null | grep maxsim_tc | \
        awk -F'","' -v ts=$1 -v dbg=$2 '{gsub(/"/,"",$NF); printf "ts=%s dbg=%s %s %s\n", ts, dbg, $(NF-2), $NF}'
done
