"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python scripts/summarise_launches.py launches.csv "header comment" [first_id last_id] > profiles/..._summary.txt"""
import csv, sys, re
path, title = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else None
hi = int(sys.argv[4]) if len(sys.argv) > 4 else None
hdr, rows = None, []
for r in csv.reader(open(path)):
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(d["ID"])
        if (lo is not None and i < lo) or (hi is not None and i > hi):
            continue
        ns = float(d["Metric Value"].replace(",", ""))
        if d.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        rows.append((name, ns))
agg = {}
for n, ns in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in agg.values())
print(f"# {title}")
print(f"# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache serialised launches: compare SHARES)")
print(f"# total {tot / 1e3:.0f} us over {len(rows)} launches")
print(f"{'kernel':66s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:66]:66s} {c:5d} {ns / 1e3:10.1f} {ns / 1e3 / c:9.2f} {ns / tot:6.3f}")
