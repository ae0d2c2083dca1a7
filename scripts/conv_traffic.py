"""profiles/conv_traffic.json from an ncu csv (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
on the conv kernels of one bench step): average DRAM bytes per conv3x3 fprop/dgrad launch, next to the algorithmic
bytes (input + output tensor once). Usage: python scripts/conv_traffic.py gpurun_out/x_convdram.csv"""
import csv
import json
import sys


def main(path, out):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = {}
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = r.get("Metric Unit", "")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = v * scale
    import re

    # conv3x3 fprop/dgrad launches only: the pair kernels, and conv3_res_kernel with TAPS = 9, KIND = 0 (the same kernel
    # also runs the 1x1 first-layer GEMM and the ConvTranspose2d GEMMs, which are not part of this figure)
    def is_conv3(name):
        return ("conv3_pair_kernel" in name or "conv3_res2_kernel" in name
                or re.search(r"conv3_res_kernel<\d+, \d+, \d+, \d+, 9, 0(, \d+)*>", name) is not None)

    rows = [d for d in per.values() if "dram__bytes_read.sum" in d and is_conv3(d["name"])]
    starts = [i for i, d in enumerate(rows)]
    if len(rows) > 34:  # the capture window may cover more than one step: keep the first whole step's 34 launches
        rows = rows[:34]
    n = len(rows)
    rd = sum(d["dram__bytes_read.sum"] for d in rows)
    wr = sum(d["dram__bytes_write.sum"] for d in rows)
    # algorithmic bytes of the 34 conv3x3 fprop/dgrad launches of one config-2 step: each reads its input and writes its
    # output once (bf16). fprop: 17 layers (all but inc.conv1); dgrad: the same 17 with in/out swapped = same bytes.
    B = 16
    layers = [(64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 512, 64), (512, 512, 64),
              (512, 1024, 32), (1024, 1024, 32), (1024, 512, 64), (512, 512, 64), (512, 256, 128), (256, 256, 128),
              (256, 128, 256), (128, 128, 256), (128, 64, 512), (64, 64, 512)]
    algo = sum(2 * B * s * s * (ci + co) for ci, co, s in layers) * 2
    res = {"launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": (rd + wr) / max(n, 1),
           "algorithmic_bytes_per_launch": algo / 34.0, "kernels": sorted({d["name"] for d in rows})}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "profiles/conv_traffic.json")
