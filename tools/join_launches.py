"""join an ncu gpu__time_duration launch list with the plan's launch descriptors; print per-kind and per-shape tables"""
import csv, sys, collections
csvf, descf = sys.argv[1], sys.argv[2]
rows = []
with open(csvf) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"], us))
descs = [l.rstrip("\n").split("\t") for l in open(descf)]
print(f"{len(rows)} kernels, {len(descs)} descriptors, total {sum(u for _, u in rows)/1000:.3f} ms")
assert len(rows) == len(descs), "launch list and plan disagree"
kind = collections.OrderedDict()
shape = {}
for (name, us), (i, k, fl, d) in zip(rows, descs):
    e = kind.setdefault(k, [0, 0.0]); e[0] += 1; e[1] += us
    s = shape.setdefault((k, d), [0, 0.0, 0.0]); s[0] += 1; s[1] += us; s[2] += float(fl)
tot = sum(u for _, u in rows)
for k, (n, us) in sorted(kind.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:20s} {n:4d} launches {us/1000:8.3f} ms {100*us/tot:5.1f} %")
print()
for (k, d), (n, us, fl) in sorted(shape.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[3]) if len(sys.argv) > 3 else 60]:
    tf = fl / (us * 1e-6) / 1e12 if us > 0 else 0
    print(f"{k:10s} x{n:3d} {us/1000:8.3f} ms {us/n:8.1f} us/launch {tf:7.1f} TF/s  {d}")
