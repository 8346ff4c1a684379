#!/bin/bash
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; env "$@" timeout 300 python tools/scratch/debug_graph2.py 2>&1 | grep -E "^(eager|graph)" | cut -c1-420; }
run DBG_WARM=1
run DBG_WARM=1 MTD_PDL=0
run DBG_WARM=1 MTDGAN_WGRAD_STREAM=0
run DBG_WARM=1 DBG_PRESERVE=0
run DBG_WARM=1 DBG_NOREPACK=1
run DBG_WARM=2
