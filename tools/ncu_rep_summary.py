"""ncu-rep -> JSON summary (per captured launch + stall mix + top source lines of one launch).
usage: python tools/ncu_rep_summary.py <rep> <out.json> "<note>" [mangled kernel name for the source page]"""
import csv, json, subprocess, sys, collections

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
cols = {"kernel": "Kernel Name", "grid": "launch__grid_size", "block": "launch__block_size", "regs": "launch__registers_per_thread",
        "smem_dyn_kb": "launch__shared_mem_per_block_dynamic", "duration_us": "gpu__time_duration.sum",
        "dram_read_mb": "dram__bytes_read.sum", "dram_write_mb": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "inst": "smsp__inst_executed.sum",
        "occ_limit_smem": "launch__occupancy_limit_shared_mem", "occ_limit_regs": "launch__occupancy_limit_registers"}
idx = {k: hdr.index(v) for k, v in cols.items() if v in hdr}
launches = []
for r in rows[2:]:
    d = {}
    for k, i in idx.items():
        v = r[i]
        if k == "kernel":
            d[k] = v.split("(")[0].split("::")[-1]
        else:
            try:
                d[k] = float(v.replace(",", ""))
            except ValueError:
                d[k] = v
    launches.append(d)
res = {"note": note, "launches": launches}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
if starts:
    sh = srows[starts[0] + 1]
    seg = srows[starts[0] + 2:(starts[1] if len(starts) > 1 else len(srows))]
    st = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    c = collections.Counter()
    for r in seg:
        for h in st:
            v = r[sh.index(h)]
            if v:
                c[h] += int(v)
    tot = sum(c.values()) or 1
    res["stall_mix_launch0"] = {k: round(v / tot, 3) for k, v in c.most_common(8)}
    isrc, isamp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
    top = sorted(seg, key=lambda r: -int(r[isamp]))[:10]
    tots = sum(int(r[isamp]) for r in seg) or 1
    res["top_sass_launch0"] = [{"sass": r[isrc].strip()[:70], "samples_pct": round(100 * int(r[isamp]) / tots, 1), "executed": int(r[iex])} for r in top]
json.dump(res, open(out, "w"), indent=1)
print(out, len(launches), "launches")
