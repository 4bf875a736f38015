#!/usr/bin/env python
"""ncu `--metrics gpu__time_duration.sum` launch list (csv) -> per-kernel totals for ONE bench step.

usage: python tools/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/rNN_launches.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
    # a step starts with the xyz kNN graph: its first kernel is the per-cloud preparation of xyz clouds (older launch
    # lists: the squared norms, three kNN graphs per step)
    starts = [i for i, n in enumerate(names) if "tcp_cloud_prep_kernel<1>" in n or "tcp_cloud_prep_kernel<true>" in n]
    if len(starts) > step + 1:
        a, b = starts[step], starts[step + 1]
    else:
        starts = [i for i, n in enumerate(names) if "sqnorm" in n]
        a, b = starts[3 * step], starts[3 * (step + 1)]
    agg = collections.OrderedDict()
    for n, v in zip(names[a:b], vals[a:b]):
        key = re.sub(r"\(.*", "", n).replace("void ", "")
        c, t = agg.get(key, (0, 0.0))
        agg[key] = (c + 1, t + v)
    tot = sum(t for _, t in agg.values())
    print(f"# kernels of bench step {step}: {b - a} launches, {tot / 1e6:.3f} ms summed (ncu, serialised, cold cache)\n")
    print("| ms | launches | share | kernel |\n|---:|---:|---:|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {t / 1e6:.3f} | {c} | {100 * t / tot:.1f}% | `{k[:110]}` |")


if __name__ == "__main__":
    main()
