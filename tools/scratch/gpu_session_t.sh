#!/bin/bash
# stream-K / tile sweep of the streamed-weight halo kernel over the eligible discriminator layer shapes
set -u
O=gpurun_out; mkdir -p $O
for B in 20 40; do
  : > $O/r2t_tune_halo_b$B.txt
  for i in 2 3 4 5 6 7 8 9 10 11; do
    timeout 300 python tools/tune_tc.py --versions 1 --passes 3 --batch $B --only $i >> $O/r2t_tune_halo_b$B.txt 2>&1
  done
  grep -c "overall best" $O/r2t_tune_halo_b$B.txt
done
