"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean and total time.
usage: launch_summary.py <csv> [skip_first_n_launches]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[1 + skip:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[hdr.index("Metric Unit")]
    v = v / 1000.0 if u in ("ns", "nsecond") else v * 1000.0 if u in ("ms", "msecond") else v
    name = r[ki].split("(")[0].replace("void ", "").replace("vct::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'n':>4s} {'mean us':>10s} {'total us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:4d} {t / n:10.2f} {t:10.1f} {100 * t / tot:6.1f}%")
print(f"{'total':60s} {'':4s} {'':10s} {tot:10.1f}")
