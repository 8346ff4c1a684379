#!/bin/bash
# Full ncu detail of the general tcgen05 conv kernel inside the train step (raw metric dump), plus a bench line.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 ncu --clock-control none --set full --import-source on --profile-from-start off -k regex:conv_tc_kernel -c 8 -f -o /tmp/ncu_tc \
    python tools/profile_step.py > $O/r2g_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu_tc.ncu-rep --page raw --csv > $O/r2g_conv_tc_raw.csv 2>/dev/null
ncu -i /tmp/ncu_tc.ncu-rep --page details > $O/r2g_conv_tc_details.txt 2>/dev/null
rm -f /tmp/ncu_tc.ncu-rep
MTD_BENCH_PER_ENTRY=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > $O/r2g_bench.json 2> $O/r2g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print("ms/step", d["ms_per_step"], "patches/s", d["value"], "launches", d.get("gpu_launches"))
kb=d["kernel_breakdown_ms"]; print({k:v for k,v in kb.items() if k not in ("per_entry","timing")})
PY
du -sh $O
