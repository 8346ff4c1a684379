#!/bin/bash
# streamed-weight halo kernel: kernel tests, microbench, bench A/B
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "streamed_halo or tcgen05_forward or halo_tile" > $O/r2h_tests.log 2>&1; echo "ktests rc=$?"; tail -15 $O/r2h_tests.log | cut -c1-300
timeout 600 python tools/bench_halo.py > $O/r2h_bench_halo.txt 2>&1; echo "bench_halo rc=$?"; cat $O/r2h_bench_halo.txt | cut -c1-200
MTD_BENCH_PER_ENTRY=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > $O/r2h_bench.json 2> $O/r2h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print("ms/step", d["ms_per_step"], "patches/s", d["value"], "launches", d.get("gpu_launches"))
kb=d["kernel_breakdown_ms"]; print({k:v for k,v in kb.items() if k not in ("per_entry","timing")})
pe=kb.get("per_entry",{}); print({k:v for k,v in pe.items() if "conv_fwd" in k or "dgrad" in k})
PY
