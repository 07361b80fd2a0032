#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and
shares; times are cold-cache and serialised, so only the SHARES are meaningful.

    python tools/launch_summary.py gpurun_out/launches.csv profiles/r1_launches.txt [first_kernel_substr]
"""
import collections
import csv
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    anchor = sys.argv[3] if len(sys.argv) > 3 else "project_cull_fused"
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            h, start = r, i + 1
            break
    iK, iV, iN = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    seq = [(r[iK], float(r[iV].replace(",", "")) / 1000.0) for r in rows[start:]
           if len(r) > iV and r[iN] == "gpu__time_duration.sum"]
    idx = [i for i, (n, _) in enumerate(seq) if anchor in n]
    lines = [f"# {src}: {len(seq)} launches; steps start at '{anchor}' ({len(idx)} found); times in us, cold-cache"]
    # the last complete fwd+bwd step: the last anchor-to-anchor window that contains a backward kernel
    win = None
    for a, b in zip(idx[:-1], idx[1:]):
        if any("bwd" in n or "backward" in n for n, _ in seq[a:b]):
            win = (a, b)
    if win is None and len(idx) >= 2:
        win = (idx[-2], idx[-1])
    if win:
        a, b = win
        tot = sum(t for _, t in seq[a:b])
        agg = collections.OrderedDict()
        for n, t in seq[a:b]:
            key = n.split("(")[0].replace("void ", "")
            c = agg.setdefault(key, [0, 0.0])
            c[0] += 1
            c[1] += t
        lines.append(f"## one fwd+bwd step: launches {a}..{b - 1} ({b - a} launches), sum {tot:.1f} us")
        lines.append(f"{'share':>7} {'us':>9} {'n':>3}  kernel")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{t / tot * 100:6.1f}% {t:9.1f} {c:3d}  {k[:110]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
