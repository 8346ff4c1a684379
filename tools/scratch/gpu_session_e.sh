#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -s -x -k "fft or generator or inference or full_train_step or b20" > $O/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2e_tests.log | cut -c1-300; grep -n "rel err\|per-slice" $O/r2e_tests.log | cut -c1-160
MTD_BENCH_PER_ENTRY=1 timeout 1200 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > $O/r2e_bench.json 2> $O/r2e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print("ms/step", d["ms_per_step"], "patches/s", d["value"])
kb=d["kernel_breakdown_ms"]; print({k:v for k,v in kb.items() if k not in ("per_entry","timing")})
print({k:v for k,v in kb.get("per_entry",{}).items() if "fft" in k or "wgrad" in k})
i=d["inference"]; print("infer b1", i["batch1"]["ms_per_slice"], i["batch1"]["kernel_ms_per_entry"], "batched", i["batched"]["value"], i["batched"]["config"][-40:])
PY
