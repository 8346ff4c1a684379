#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -s -x \
   -k "fft or halo or generator or tcgen05_forward" > $O/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/r2c_tests.log | cut -c1-300
timeout 300 python tools/scratch/debug_graph.py 1 > $O/r2c_dbg_default.log 2>&1; grep trial $O/r2c_dbg_default.log | cut -c1-400
MTD_PDL=0 timeout 300 python tools/scratch/debug_graph.py 1 > $O/r2c_dbg_nopdl.log 2>&1; grep trial $O/r2c_dbg_nopdl.log | cut -c1-400
MTDGAN_WGRAD_STREAM=0 timeout 300 python tools/scratch/debug_graph.py 1 > $O/r2c_dbg_nows.log 2>&1; grep trial $O/r2c_dbg_nows.log | cut -c1-400
timeout 300 python tools/scratch/debug_graph.py 2 > $O/r2c_dbg_warm2.log 2>&1; grep trial $O/r2c_dbg_warm2.log | cut -c1-400
timeout 600 python tools/bench_c32.py > $O/r2c_bench_c32.txt 2>&1; echo "c32 rc=$?"; grep "passes=3 halo\|passes=1 halo" $O/r2c_bench_c32.txt
MTD_BENCH_PER_ENTRY=1 timeout 1200 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print("ms/step", d["ms_per_step"], "patches/s", d["value"])
kb=d["kernel_breakdown_ms"]; print({k:v for k,v in kb.items() if k not in ("per_entry","timing")})
print({k:v for k,v in kb.get("per_entry",{}).items() if "fft" in k})
i=d["inference"]; print("infer b1", i["batch1"]["ms_per_slice"], i["batch1"]["kernel_ms_per_entry"], "batched", i["batched"]["value"], i["batched"]["config"][-40:])
PY
du -sh $O
