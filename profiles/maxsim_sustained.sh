# sustained single-query MaxSim under the power cap: operand placement x timing-only debug variants
for cfg in "0 0" "0 32" "0 16" "0 64" "0 4" "0 2" "0 1" "1 0" "1 8" "1 24" "1 32" "1 64"; do set -- $cfg
  echo -n "ts=$1 dbg=$2: "; INNR_MAXSIM_TS=$1 INNR_MAXSIM_DEBUG=$2 timeout 120 python innr_b200/csrc/dev/maxsim_sustained.py 2>&1 | tail -1
done
echo -n "pair ts=0: "; INNR_MAXSIM_TS=0 timeout 120 python innr_b200/csrc/dev/maxsim_sustained.py pair 2>&1 | tail -1
echo -n "pair ts=1: "; INNR_MAXSIM_TS=1 timeout 120 python innr_b200/csrc/dev/maxsim_sustained.py pair 2>&1 | tail -1
