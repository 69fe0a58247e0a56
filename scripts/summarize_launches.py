"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rd:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"^void ", "", name); name = re.sub(r"\(anonymous namespace\)::", "", name)
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    c, t = agg.get(name, (0, 0.0)); agg[name] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
print(f"# {path}: {sum(c for c,_ in agg.values())} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':70s} {'count':>6s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} {c:6d} {t:12.1f} {t/c:10.2f} {100*t/tot:6.1f}%")
