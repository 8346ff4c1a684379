#!/usr/bin/env python
"""Key-metric summary of an `ncu --set full` report, one row per kernel name (mean over its captured launches):
duration, DRAM bytes and % of peak, L2 throughput %, tensor-pipe %, shared-memory bank conflicts, occupancy, registers.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rNN_ncu_<what>.txt]   (runs `ncu -i ... --page raw --csv`)"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),                      # ns -> us
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1),
    ("lts__t_bytes.sum", "l2_MB", 1e-6),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 1),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
              "ns": 1.0, "us": 1e3, "ms": 1e6}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols, units = rows[hdr], rows[hdr + 1]
    ki = cols.index("Kernel Name")
    idx = {}
    for name, short, sc in WANT:
        if name in cols:
            idx[short] = (cols.index(name), sc)
    agg = collections.OrderedDict()
    for r in rows[hdr + 2:]:
        if len(r) <= ki:
            continue
        k = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void\s+", "", r[ki].split("(")[0])
        d = agg.setdefault(k, collections.defaultdict(list))
        for short, (ci, sc) in idx.items():
            try:
                v = float(r[ci].replace(",", ""))
            except ValueError:
                continue
            v *= UNIT_SCALE.get(units[ci], 1.0) if short in ("dur_us", "dram_rd_MB", "dram_wr_MB", "l2_MB") else 1.0
            d[short].append(v * sc)
    shorts = [s for _, s, _ in WANT if s in idx]
    print(f"# {path}: mean over captured launches per kernel (ncu --set full, --clock-control none; cold-cache, serialised)")
    print(f"{'kernel':58s} {'n':>3s} " + " ".join(f"{s:>14s}" for s in shorts))
    for k, d in agg.items():
        n = max(len(v) for v in d.values())
        print(f"{k[:58]:58s} {n:3d} " + " ".join(f"{(sum(d[s]) / len(d[s]) if d[s] else float('nan')):14.3f}" for s in shorts))


if __name__ == "__main__":
    main(sys.argv[1])
