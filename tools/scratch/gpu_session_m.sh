#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tcgen05 or halo or streamed or stride2" > $O/r2m_tests.log 2>&1; echo "ktests rc=$?"; tail -4 $O/r2m_tests.log | cut -c1-300
MTD_BENCH_PER_ENTRY=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > $O/r2m_bench.json 2> $O/r2m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
print("ms/step", d["ms_per_step"], "patches/s", d["value"], "launches", d.get("gpu_launches"))
kb=d["kernel_breakdown_ms"]; print({k:v for k,v in kb.items() if k not in ("per_entry","timing")})
pe=kb.get("per_entry",{}); print({k:v for k,v in pe.items() if "conv_fwd" in k or "dgrad" in k})
PY
