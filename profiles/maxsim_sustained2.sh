for cfg in "1 0" "1 128" "1 256" "1 384" "1 440"; do set -- $cfg
  echo -n "ts=$1 dbg=$2: "; INNR_MAXSIM_TS=$1 INNR_MAXSIM_DEBUG=$2 timeout 120 python innr_b200/csrc/dev/maxsim_sustained.py 2>&1 | tail -1
done
