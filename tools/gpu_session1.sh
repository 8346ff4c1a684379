#!/bin/bash
# Round-2 GPU session 1: smoke, full GPU test suite (with gradient-error tallies), bench line, launch lists and
# `ncu --set full` captures of every kernel family INSIDE the train step / the 512x512 inference forward.
# Usage (from the repo root, under gpurun): bash tools/gpu_session1.sh [tests|bench|ncu ...]   (default: all)
set -u
O=gpurun_out
mkdir -p $O
WHAT="${*:-smoke tests bench launches ncu}"
NCU="ncu --clock-control none"
for w in $WHAT; do
case $w in
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" ;;
tests)
  rm -f $O/r2_tally.jsonl
  MTD_TALLY_DUMP=$O/r2_tally.jsonl timeout 2400 python -m pytest tests -m gpu -x -q -s > $O/r2_tests.log 2>&1; echo "tests rc=$?"
  tail -5 $O/r2_tests.log ;;
bench)
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt
  nproc > $O/r2_nproc.txt
  MTD_BENCH_PER_ENTRY=1 timeout 1500 python bench.py --steps 10 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err; echo "bench rc=$?"
  head -c 600 $O/r2_bench.json; echo ;;
launches)
  timeout 900 $NCU --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --csv --log-file $O/r2_launches_train.csv \
      python tools/profile_step.py > $O/r2_launches_train.log 2>&1; echo "launches rc=$?" ;;
ncu)
  # one invocation per kernel family so every family gets its own launch budget (-c); all inside the NVTX range of
  # ONE steady-state train step (eager launches, B = 20)
  i=0
  for spec in "conv_tc_kernel:8" "wgrad_tc_kernel:6" "fft_:12" "sn_:8" "pcgrad:4" "adamw:3" "conv_c1|conv_n1|thin_wgrad:9" \
              "finish:8" "act_bwd:6" "conv_igemm|conv_wgrad_kernel:6" "upsample|pixel_shuffle|edge|sum_|loss:10"; do
    pat="${spec%%:*}"; cnt="${spec##*:}"; i=$((i+1))
    timeout 600 $NCU --set full --import-source on --nvtx --nvtx-include "timed/" -k "regex:$pat" -c $cnt -f -o $O/r2_ncu_train_$i \
        python tools/profile_step.py > $O/r2_ncu_train_$i.log 2>&1; echo "ncu train [$pat] rc=$?"
  done
  timeout 900 $NCU --set full --nvtx --nvtx-include "timed/" -f -o $O/r2_ncu_infer512 \
      python tools/profile_step.py --what infer > $O/r2_ncu_infer512.log 2>&1; echo "ncu infer rc=$?" ;;
esac
done
ls -la $O | grep r2_ | head -40
