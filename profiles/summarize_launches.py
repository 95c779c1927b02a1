"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel summary
(markdown).  Usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/xxx.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import io
import re
import sys


def load(path):
    txt = open(path).read()
    lines = [l for l in txt.splitlines() if l.startswith('"')]
    rd = csv.reader(io.StringIO("\n".join(lines)))
    hdr = next(rd)
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    out = []
    for r in rd:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki]
        m = re.search(r"([A-Za-z_0-9]+_kernel(?:<[^>]*>)?)", name)
        short = m.group(1) if m else name.split("(")[0][:60]
        out.append((short, r[gi], r[bi], v / 1e3))
    return out


def main():
    seq = load(sys.argv[1])
    title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    agg = collections.OrderedDict()
    for name, grid, block, us in seq:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] = max(a[2], us)
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {title}\n")
    print(f"{len(seq)} launches, {tot:.1f} us summed device time (serialised under ncu; see the file name for the cache control used).\n")
    print("| kernel | launches | total us | avg us | max us | share |\n|---|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.2f} | {a[2]:.2f} | {100 * a[1] / tot:.1f}% |")
    # one GN iteration in launch order
    try:
        starts = [i for i, s in enumerate(seq) if s[0].startswith("assemble_pose")]
        a, b = starts[1], starts[2]
        print("\n## one Gauss-Newton iteration, launch order\n")
        print("| kernel | grid | block | us |\n|---|---|---|---:|")
        for name, grid, block, us in seq[a:b]:
            print(f"| `{name}` | {grid} | {block} | {us:.2f} |")
        print(f"\nsum = {sum(s[3] for s in seq[a:b]):.1f} us")
    except Exception:
        pass


if __name__ == "__main__":
    main()
