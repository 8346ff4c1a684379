#!/bin/bash
# Full ncu detail of the streamed-weight halo kernel inside the train step (raw metric dump).
set -u
O=gpurun_out; mkdir -p $O
timeout 900 ncu --clock-control none --set full --import-source on --profile-from-start off -k regex:conv_halo_kernel -c 10 -f -o /tmp/ncu_halo \
    python tools/profile_step.py > $O/r2q_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu_halo.ncu-rep --page raw --csv > $O/r2q_conv_halo_raw.csv 2>/dev/null
ncu -i /tmp/ncu_halo.ncu-rep --page details > $O/r2q_conv_halo_details.txt 2>/dev/null
ncu -i /tmp/ncu_halo.ncu-rep --page source --csv --kernel-name regex:conv_halo_kernel --launch-skip 1 --launch-count 1 > $O/r2q_conv_halo_source.csv 2>/dev/null
ls -la /tmp/ncu_halo.ncu-rep
rm -f /tmp/ncu_halo.ncu-rep
du -sh $O
