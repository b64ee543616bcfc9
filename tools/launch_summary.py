"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (development aid)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except Exception:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(d["Metric Unit"], 1.0)
    name = d["Kernel Name"].split("(")[0][-70:]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(t for _, t in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t:12.1f} us {100 * t / tot:5.1f}%  {n:6d} x {t / n:10.1f} us  {k}")
