"""join an ncu launch list (gpu__time_duration [+ dram bytes]) with the plan's launch descriptors; per-kind / per-shape tables"""
import csv, sys, collections
csvf, descf = sys.argv[1], sys.argv[2]
traffic_json = None
if "--traffic-json" in sys.argv:  # write {gemm_launches, dram_bytes, ...} for bench.py's roofline.traffic
    i = sys.argv.index("--traffic-json")
    traffic_json = sys.argv[i + 1]
    del sys.argv[i:i + 2]
with open(csvf) as f:
    lines = [l for l in f if not l.startswith("==")]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    e = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        e["us"] = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    elif m.startswith("dram__bytes"):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        e[m] = v * scale
rows = list(per.values())
descs = [l.rstrip("\n").split("\t") for l in open(descf)]
print(f"{len(rows)} kernels, {len(descs)} descriptors, total {sum(r['us'] for r in rows)/1000:.3f} ms (cold-cache, serialised ncu timings)")
assert len(rows) == len(descs), "launch list and plan disagree"
kind, shape = collections.OrderedDict(), {}
for r, (i, k, fl, d) in zip(rows, descs):
    dram = r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
    e = kind.setdefault(k, [0, 0.0, 0.0, 0.0]); e[0] += 1; e[1] += r["us"]; e[2] += dram; e[3] += float(fl)
    s = shape.setdefault((k, d), [0, 0.0, 0.0, 0.0]); s[0] += 1; s[1] += r["us"]; s[2] += float(fl); s[3] += dram
tot = sum(r["us"] for r in rows)
if traffic_json:
    import json
    n, us, dram, fl = kind["gemm"]
    with open(traffic_json, "w") as f:
        json.dump({"gemm_launches": n, "dram_bytes": dram, "gemm_us_cold_serialised": us, "plan_launches": len(rows),
                   "source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, {csvf.split('/')[-1]} "
                             "(one eager forward of the plan bench.py times; cold-cache, serialised)"}, f)
for k, (n, us, dram, fl) in sorted(kind.items(), key=lambda kv: -kv[1][1]):
    extra = f" {fl/(us*1e-6)/1e12:7.1f} TFLOP/s" if fl > 0 else f" {dram/(us*1e-6)/1e9:7.0f} GB/s DRAM"
    print(f"{k:20s} {n:4d} launches {us/1000:8.3f} ms {100*us/tot:5.1f} %  dram {dram/1e6:9.1f} MB{extra}")
print()
for (k, d), (n, us, fl, dram) in sorted(shape.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[3]) if len(sys.argv) > 3 else 60]:
    tf = fl / (us * 1e-6) / 1e12 if us > 0 else 0
    print(f"{k:10s} x{n:3d} {us/1000:8.3f} ms {us/n:8.1f} us/launch {tf:7.1f} TF/s dram {dram/n/1e6:7.1f} MB/launch  {d}")
