"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` log per kernel.

usage: python tools/ncu_launch_list.py <log.csv> <out.json> "<command note>" [first_id last_id]
Launch ids select one training step (the log holds warm-up + timed steps)."""
import csv, json, re, sys, collections

log, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 1 << 60
rows = [r for r in csv.reader(l for l in open(log) if l.startswith('"'))]
hdr = rows[0]
iid, ik, im, iv, iu = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
launch = collections.OrderedDict()
for r in rows[1:]:
    i = int(r[iid])
    if i < lo or i > hi:
        continue
    d = launch.setdefault(i, {"kernel": r[ik]})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[im]] = v * scale
agg = collections.OrderedDict()
for d in launch.values():
    name = re.sub(r"\(.*$", "", d["kernel"])
    name = name.split("::")[-1] if "<" not in name else name.replace("dsg::", "").replace("tc::", "")
    a = agg.setdefault(name, {"kernel": name, "launches": 0, "us": 0.0, "dram_read_mb": 0.0, "dram_write_mb": 0.0})
    a["launches"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    a["dram_read_mb"] += d.get("dram__bytes_read.sum", 0.0) / 1e6
    a["dram_write_mb"] += d.get("dram__bytes_write.sum", 0.0) / 1e6
total = sum(a["us"] for a in agg.values())
ks = sorted(agg.values(), key=lambda a: -a["us"])
for a in ks:
    a["share"] = round(a["us"] / total, 4)
    a["us"] = round(a["us"], 1)
    a["dram_read_mb"] = round(a["dram_read_mb"], 1)
    a["dram_write_mb"] = round(a["dram_write_mb"], 1)
    a["dram_gbs"] = round((a["dram_read_mb"] + a["dram_write_mb"]) * 1e6 / (a["us"] * 1e-6) / 1e9, 1) if a["us"] else 0.0
json.dump({"command": note, "launch_ids": [lo, min(hi, max(launch) if launch else 0)], "launches": len(launch), "total_us": round(total, 1),
           "kernels": ks}, open(out, "w"), indent=1)
print(f"{len(launch)} launches, {total:.1f} us")
for a in ks[:14]:
    print(f"{a['kernel'][:44]:44s} n={a['launches']:4d} us={a['us']:9.1f} share={a['share']:.3f} dram={a['dram_read_mb'] + a['dram_write_mb']:8.1f} MB  {a['dram_gbs']:7.1f} GB/s")
