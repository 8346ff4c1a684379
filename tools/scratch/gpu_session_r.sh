#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 ncu --clock-control none --set full --profile-from-start off -k regex:conv_tc_kernel -c 8 -f -o /tmp/ncu_tc \
    python tools/profile_step.py > $O/r2r_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu_tc.ncu-rep --page raw --csv > $O/r2r_conv_tc_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/ncu_tc.ncu-rep > $O/r2r_conv_tc_summary.txt 2>&1
rm -f /tmp/ncu_tc.ncu-rep
python tools/ncu_limiter.py $O/r2r_conv_tc_raw.csv conv_tc_kernel
