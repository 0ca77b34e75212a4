"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections, csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict(); n = 0
for row in csv.DictReader(io.StringIO("".join(lines))):
    n += 1
    a = agg.setdefault(row["Kernel Name"][:110], [0, 0.0]); a[0] += 1; a[1] += float(row["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot/1e6:.3f} ms total (serialised, cold-cache: compare SHARES)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t/1e3:10.1f} us {100*t/tot:5.1f}%  x{c:<4d} avg {t/c/1e3:8.1f} us  {k}")
