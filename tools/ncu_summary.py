"""Markdown summary of an `ncu --set full` report (one row per profiled launch).

    python tools/ncu_summary.py gpurun_out/r02_kernels.ncu-rep [label ...] > profiles/r02_kernels_ncu_table.md

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU) and prints, per launch: duration, grid x block,
registers, DRAM bytes read + written, DRAM / L2 / L1TEX throughput (% of peak), L2 hit rate, tensor-pipe and issue activity,
achieved occupancy and the top warp-stall reasons.  Optional labels name the launches in order."""
from __future__ import annotations

import csv
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time us", lambda v, u: "%.1f" % (v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u == "ms" else 1.0))),
    ("launch__grid_size", "grid", lambda v, u: "%d" % v),
    ("launch__block_size", "block", lambda v, u: "%d" % v),
    ("launch__registers_per_thread", "regs", lambda v, u: "%d" % v),
    ("dram__bytes_read.sum", "DRAM rd MB", None),
    ("dram__bytes_write.sum", "DRAM wr MB", None),
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "DRAM busy %", lambda v, u: "%.1f" % v),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", lambda v, u: "%.1f" % v),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", lambda v, u: "%.1f" % v),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX %", lambda v, u: "%.1f" % v),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)", lambda v, u: "%.1f" % v),
    ("sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed", "tf32 ops % of peak", lambda v, u: "%.1f" % v),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %", lambda v, u: "%.1f" % v),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %", lambda v, u: "%.1f" % v),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", lambda v, u: "%.1f" % v),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", lambda v, u: "%.1f" % v),
]
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rep = sys.argv[1]
    labels = sys.argv[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    kname = ix.get("Kernel Name")
    stall_cols = [(h, i) for h, i in ix.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                  or h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
    head = ["#", "kernel"] + [c[1] for c in COLS] + ["top stalls (warp cycles per issue)"]
    print("| " + " | ".join(head) + " |")
    print("|" + "---|" * len(head))
    for n, r in enumerate(data):
        name = r[kname].split("(")[0] if kname is not None else "?"
        if n < len(labels):
            name += " -- " + labels[n]
        cells = [str(n), name]
        for key, _, fmt in COLS:
            i = ix.get(key)
            v = num(r[i]) if i is not None else None
            if v is None:
                cells.append("n/a")
            elif fmt is None:
                cells.append("%.2f" % (v * BYTES.get(units[i], 1.0) / 1e6))
            else:
                cells.append(fmt(v, units[i]))
        st = []
        for h, i in stall_cols:
            v = num(r[i])
            if v:
                st.append((v, h.split("issue_stalled_")[1].rsplit("_per_issue", 1)[0].replace(".ratio", "")))
        st.sort(reverse=True)
        cells.append(", ".join("%s %.1f" % (k, v) for v, k in st[:3]))
        print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
