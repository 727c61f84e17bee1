# load-path experiments (TMA streaming only, no compute) at locked base clocks: box rows / stream order / L2 promotion
for cfg in "0 128" "32 128" "64 128" "128 128" "0 256" "128 256" "0 0"; do set -- $cfg
  INNR_MAXSIM_TS=0 INNR_MAXSIM_DEBUG=1 INNR_MAXSIM_BOXROWS=$1 INNR_MAXSIM_L2PROMO=$2 ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,dram__bytes_read.sum --clock-control base \
        -k regex:maxsim_tc_kernel -s 2 -c 1 --csv python innr_b200/csrc/dev/maxsim_prof.py 2>/dev/null | grep maxsim_tc | \
        awk -F'","' -v a=$1 -v b=$2 '{gsub(/"/,"",$NF); printf "boxrows=%s promo=%s %s %s\n", a, b, $(NF-2), $NF}'
done
