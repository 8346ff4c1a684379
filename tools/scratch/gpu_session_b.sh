#!/bin/bash
# short GPU session: targeted tests, lr debug, c32 microbenchmark, ncu of the c32 kernel, full-step launch list
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ablation.py tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -s \
   -k "ablation or halo or full_train_step_b4 or lr_schedule or graph_step_matches" > $O/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/r2b_tests.log
timeout 600 python tools/scratch/debug_lr.py > $O/r2b_debug_lr.log 2>&1; echo "debug rc=$?"; grep -E "^(eager|graph)" $O/r2b_debug_lr.log
timeout 600 python tools/bench_c32.py > $O/r2b_bench_c32.txt 2>&1; echo "c32 rc=$?"; cat $O/r2b_bench_c32.txt
timeout 600 ncu --clock-control none --set full --import-source on --profile-from-start off -k "regex:conv_c32" -c 3 -f -o $O/r2b_ncu_c32 \
   python tools/profile_step.py --what infer > $O/r2b_ncu_c32.log 2>&1; echo "ncu c32 rc=$?"
python tools/ncu_summary.py $O/r2b_ncu_c32.ncu-rep > $O/r2b_ncu_c32.txt 2>&1; cat $O/r2b_ncu_c32.txt
timeout 900 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r2b_launches_train.csv \
   python tools/profile_step.py > $O/r2b_launches_train.log 2>&1; echo "launches rc=$?"
python tools/launch_summary.py $O/r2b_launches_train.csv 70 > $O/r2b_launches_train.txt 2>&1; head -50 $O/r2b_launches_train.txt
du -sh $O
