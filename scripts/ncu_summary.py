"""Summarise an ncu report: one row per captured launch with the metrics DESIGN.md quotes.

    python scripts/ncu_summary.py gpurun_out/X.ncu-rep [--csv out.csv]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__inst_executed.sum", "warp_inst_M", 1e-6),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes", 1),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_pct", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct", 1),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_wavefront_pct", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts_M", 1e-6),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__shared_mem_per_block_static", "smem_B", 1),
    ("launch__grid_size", "grid", 1),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    table = []
    for r in rows[2:]:
        rec = {"kernel": r[col["Kernel Name"]][:48]}
        for m, name, scale in METRICS:
            if m in col:
                try:
                    v = float(r[col[m]].replace(",", ""))
                    u = units[col[m]]
                    if name == "time_us":
                        v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                    elif name.endswith("_MB"):
                        v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
                    else:
                        v = v * scale
                    rec[name] = round(v, 3)
                except ValueError:
                    rec[name] = r[col[m]]
        table.append(rec)
    keys = ["kernel"] + [n for _, n, _ in METRICS]
    if "--csv" in sys.argv:
        w = csv.DictWriter(open(sys.argv[sys.argv.index("--csv") + 1], "w"), fieldnames=keys)
        w.writeheader()
        for rec in table:
            w.writerow(rec)
    for rec in table:   # transposed, readable
        print("== " + rec["kernel"])
        print("   " + "  ".join(f"{k}={rec.get(k)}" for k in keys[1:]))


if __name__ == "__main__":
    main()
