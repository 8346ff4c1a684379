#!/usr/bin/env python
"""Per-launch limiter table from an `ncu -i report --page raw --csv` dump of tensor-core conv kernels: duration, tensor
pipe, the two consumers of the shared-memory data pipe (tensor-core operand reads, LSU traffic), L2 -> L1 and LTS shares.

    python tools/ncu_limiter.py gpurun_out/r2q_conv_halo_raw.csv conv_halo_kernel > profiles/...txt"""
import csv
import sys

COLS = [
    ("dur_us", "gpu__time_duration.sum", 1.0),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("TCsmem%", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("LSUsmem%", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("ld%", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("st%", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("xbar>L1%", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("LTS%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("SM%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("dramrdMB", "dram__bytes_read.sum", 1.0),
    ("grid", "launch__grid_size", 1.0),
]


def main():
    path, pat = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full --clock-control none, launches INSIDE one steady-state train step; pct = % of peak sustained over elapsed cycles")
    print(f"{'kernel':30s} " + " ".join(f"{c[0]:>9s}" for c in COLS))
    for r in data:
        name = r[ix["Kernel Name"]]
        if pat not in name:
            continue
        short = name[name.index(pat):].split("(")[0]
        vals = []
        for label, key, _ in COLS:
            try:
                v = float(r[ix[key]].replace(",", ""))
                if label == "dramrdMB" and units[ix[key]].lower().startswith("byte"):
                    v /= 1e6
                elif label == "dramrdMB" and units[ix[key]].lower().startswith("kbyte"):
                    v /= 1e3
            except (KeyError, ValueError):
                v = float("nan")
            vals.append(v)
        print(f"{short:30s} " + " ".join(f"{v:9.1f}" for v in vals))


if __name__ == "__main__":
    main()
