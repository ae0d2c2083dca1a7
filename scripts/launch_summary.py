"""Aggregate an ncu launch list (gpu__time_duration.sum per launch, csv) into a per-kernel table for ONE training step:
the launches between two consecutive launches of the first-layer kernel (conv3_res_kernel<..., TAPS=1, KIND=3 (RES_FIRST), CIN>;
first_im2col_kernel in round-1 captures). Usage: python scripts/launch_summary.py x.csv"""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((r["Kernel Name"], ns))
    import re

    first = re.compile(r"first_im2col|conv3_res_kernel<\d+, \d+, \d+, \d+, 1, 3, \d+(, \w+)?>")
    starts = [i for i, (n, _) in enumerate(rows) if first.search(n)]
    if len(starts) >= 2:
        step = rows[starts[0]:starts[1]]
        note = f"one step = launches {starts[0]}..{starts[1] - 1} of the capture"
    else:
        step, note = rows, "no step boundary found: whole capture"
    agg = OrderedDict()
    for n, ns in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"{note}; {len(step)} launches, {tot / 1e6:.3f} ms summed device time (ncu: serialised, cold caches)\n")
    print("| kernel | launches | ms | share |")
    print("|---|---|---|---|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n[:100]} | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
