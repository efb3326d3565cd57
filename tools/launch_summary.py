"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file) by kernel.
usage: python tools/launch_summary.py launches.csv "<command that was profiled>" > profiles/rNN_launches.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
print(f"# ncu launch list: `{cmd}`")
print("# (cold-cache, serialised per-launch times: compare SHARES, not absolutes; includes bench.py's input synthesis by torch)")
print(f"# total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
print("count  total_ms  share  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:5d} {a[1]:10.3f} {100*a[1]/tot:6.1f}%  {k[:150]}")
