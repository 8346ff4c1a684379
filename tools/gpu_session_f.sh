#!/bin/bash
# A/B of the fused stream-K reduction: kernel tests, then the bench with the fused path on and off.
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_streamk or tcgen05 or halo or conv_forward_backward" > $O/r2f_tests.log 2>&1; echo "ktests rc=$?"; tail -5 $O/r2f_tests.log | cut -c1-300
for f in 1 0; do
MTD_SK_FUSED=$f MTD_BENCH_PER_ENTRY=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > $O/r2f_bench_$f.json 2> $O/r2f_bench_$f.err; echo "bench fused=$f rc=$?"
python - $f <<'PY'
import json, sys
d=json.load(open('gpurun_out/r2f_bench_%s.json' % sys.argv[1]))
print("ms/step", d["ms_per_step"], "patches/s", d["value"], "launches", d.get("gpu_launches"))
kb=d["kernel_breakdown_ms"]; print({k:v for k,v in kb.items() if k not in ("per_entry","timing")})
pe=kb.get("per_entry",{}); print({k:v for k,v in pe.items() if "conv_c1" in k or "conv_fwd" in k or "dgrad" in k})
i=d["inference"]; print("infer b1", i["batch1"]["ms_per_slice"], "batched", i["batched"]["value"])
PY
done
