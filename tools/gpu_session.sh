#!/bin/bash
# Round-2 GPU session: kernel tests first, then the full GPU suite (with gradient-error tallies), the bench line, a
# launch list and `ncu --set full` captures of every kernel family INSIDE the train step / the 512x512 inference forward.
# ncu reports are summarised ON THE BOX (tools/ncu_summary.py) and deleted: gpurun_out/ must stay under 64 MiB.
# Usage (repo root, under gpurun): bash tools/gpu_session.sh [ktests|tests|bench|launches|ncu|ncuinfer ...]
set -u
O=gpurun_out
mkdir -p $O
WHAT="${*:-smoke ktests tests bench launches ncu ncuinfer}"
NCU="ncu --clock-control none"
TAG="${TAG:-r2}"
for w in $WHAT; do
case $w in
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log ;;
ktests)
  timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tcgen05_forward or halo or streamed" > $O/${TAG}_ktests.log 2>&1; echo "ktests rc=$?"
  tail -4 $O/${TAG}_ktests.log ;;
tests)
  rm -f $O/${TAG}_tally.jsonl
  MTD_TALLY_DUMP=$O/${TAG}_tally.jsonl timeout 2400 python -m pytest tests -m gpu -q -s > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?"
  tail -12 $O/${TAG}_tests.log ;;
bench)
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt
  nproc > $O/${TAG}_nproc.txt
  MTD_BENCH_PER_ENTRY=1 timeout 1500 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
  head -c 400 $O/${TAG}_bench.json; echo; tail -3 $O/${TAG}_bench.err ;;
launches)
  timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_train.csv \
      python tools/profile_step.py > $O/${TAG}_launches_train.log 2>&1; echo "launches rc=$?"
  python tools/launch_summary.py $O/${TAG}_launches_train.csv 60 > $O/${TAG}_launches_train.txt 2>&1
  timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_infer.csv \
      python tools/profile_step.py --what infer > $O/${TAG}_launches_infer.log 2>&1; echo "launches infer rc=$?"
  python tools/launch_summary.py $O/${TAG}_launches_infer.csv 60 > $O/${TAG}_launches_infer.txt 2>&1 ;;
ncu)
  i=0
  for spec in "conv_tc_kernel:8" "conv_halo_kernel:8" "conv_c32_kernel:6" "wgrad_tc_kernel:6" "fft_:12" "sn_:8" "pcgrad:4" "adamw:3" "conv_c1|conv_n1|thin_wgrad:9" \
              "finish:8" "act_bwd:6" "conv_igemm|conv_wgrad_kernel:6" "upsample|pixel_shuffle|edge|sum_|loss:10"; do
    pat="${spec%%:*}"; cnt="${spec##*:}"; i=$((i+1))
    timeout 600 $NCU --set full --profile-from-start off -k "regex:$pat" -c $cnt -f -o /tmp/ncu_train_$i \
        python tools/profile_step.py > $O/${TAG}_ncu_train_$i.log 2>&1; echo "ncu train [$pat] rc=$?"
    python tools/ncu_summary.py /tmp/ncu_train_$i.ncu-rep >> $O/${TAG}_ncu_train_step.txt 2>&1
    rm -f /tmp/ncu_train_$i.ncu-rep $O/${TAG}_ncu_train_$i.log
  done ;;
ncuinfer)
  timeout 900 $NCU --set full --profile-from-start off -f -o /tmp/ncu_infer512 \
      python tools/profile_step.py --what infer > $O/${TAG}_ncu_infer512.log 2>&1; echo "ncu infer rc=$?"
  python tools/ncu_summary.py /tmp/ncu_infer512.ncu-rep > $O/${TAG}_ncu_infer512_by_kernel.txt 2>&1
  ncu -i /tmp/ncu_infer512.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      > $O/${TAG}_ncu_infer512_launches.csv 2>/dev/null
  rm -f /tmp/ncu_infer512.ncu-rep ;;
esac
done
du -sh $O
