#!/usr/bin/env python
"""Turn an `ncu --page raw --csv` export (one row per kernel launch) into the per-kernel table kept under profiles/.

    python profiles/summarise_ncu.py gpurun_out/<tag>_frame_raw.csv > profiles/<tag>_ncu_kernels.md
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return hdr.index(name) if name in hdr else None


def get(r, name, default=float("nan")):
    i = col(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def unit(name):
    i = col(name)
    return units[i] if i is not None else ""


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


print("| kernel | grid x block | regs | time us | warp instr (M) | issue-active % (while active) | issue-active % of elapsed | "
      "achieved occupancy % | DRAM read MB | DRAM write MB | L2 bytes MB | DRAM % of peak | top stall reasons |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
stall_cols = [h for h in hdr if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio")
              or h.startswith("smsp__average_warps_issue_stalled")]
for r in data:
    name = r[col("Kernel Name")].split("(")[0].replace("void ", "").replace("gsr::", "")
    t = get(r, "gpu__time_duration.sum")
    tu = unit("gpu__time_duration.sum")
    t_us = t * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(tu, 1)
    inst = get(r, "smsp__inst_executed.sum") / 1e6
    ia = get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    cyc_act, cyc_el = get(r, "smsp__cycles_active.avg"), get(r, "smsp__cycles_elapsed.avg", get(r, "sm__cycles_elapsed.max"))
    ia_el = ia * cyc_act / cyc_el if cyc_el == cyc_el and cyc_el else float("nan")
    occ = get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    dr = to_bytes(get(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum")) / 1e6
    dw = to_bytes(get(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum")) / 1e6
    l2 = to_bytes(get(r, "lts__t_bytes.sum"), unit("lts__t_bytes.sum")) / 1e6
    dpct = get(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")
    stalls = sorted(((get(r, h, 0.0), h) for h in stall_cols), reverse=True)[:3]
    st = ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.1f}"
                   for v, h in stalls if v > 0)
    print(f"| {name} | {int(get(r, 'launch__grid_size', 0))} x {int(get(r, 'launch__block_size', 0))} | "
          f"{int(get(r, 'launch__registers_per_thread', 0))} | {t_us:.1f} | {inst:.2f} | {ia:.1f} | {ia_el:.1f} | {occ:.1f} | "
          f"{dr:.1f} | {dw:.1f} | {l2:.1f} | {dpct:.1f} | {st} |")
