"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/ (run here, no GPU)."""
import csv
import subprocess
import sys

KEYS = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time_us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("smsp__inst_executed.sum", "warp_insts"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stalls = [h for h in hdr if h.startswith(STALL) and h.endswith("per_issue_active.ratio")]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([k for _, k in KEYS] + ["units:time/dram", "top_stalls(per issue)"])
        for r in rows[2:]:
            vals = []
            for h, _ in KEYS:
                v = r[hdr.index(h)] if h in hdr else ""
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    v = v[:70]
                vals.append(v)
            un = f"{units[hdr.index('gpu__time_duration.sum')]}/{units[hdr.index('dram__bytes_read.sum')]}"
            st = sorted(((float(r[hdr.index(s)] or 0), s[len(STALL):-len('_per_issue_active.ratio')]) for s in stalls), reverse=True)[:5]
            w.writerow(vals + [un, "; ".join(f"{n}={v:.2f}" for v, n in st)])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
