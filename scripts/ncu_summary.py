"""Compact per-launch summary of an `ncu --set full` report (read on the CPU box): duration, tensor-pipe activity,
DRAM / L2 traffic and throughput percentages. Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x.md]"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_%", 1),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_MB", 1e-6),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr_%", 1),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%", 1),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("lts__t_bytes.sum", "l2_MB", 1e-6),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("sm__cycles_elapsed.max", "cycles", 1),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {}
    for i, h in enumerate(hdr):
        idx.setdefault(h, i)
        idx.setdefault(h.split(".", 2)[-1] if h.startswith(("TPC.", "SM_")) else h, i)
    cols = [(m, n, s) for m, n, s in METRICS if m in idx]
    print("| # | kernel | " + " | ".join(n for _, n, _ in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, r in enumerate(data):
        name = r[idx["Kernel Name"]]
        # template arguments are lost in the short name; the demangled one carries them
        full = r[idx["Function Name"]] if "Function Name" in idx else name
        vals = []
        for m, n, s in cols:
            v = r[idx[m]].replace(",", "")
            u = units[idx[m]]
            try:
                f = float(v)
                if m == "gpu__time_duration.sum":
                    f = f * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                    vals.append(f"{f:.1f}")
                elif u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
                    vals.append(f"{f:.1f}")
                else:
                    vals.append(f"{f:.1f}" if f != int(f) else str(int(f)))
            except ValueError:
                vals.append(v)
        print(f"| {k} | {full[:90]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
