"""Turns gpurun_out/{launches_<tag>.csv, prof_<tag>.ncu-rep, bench_<tag>.json} into tracked summaries under profiles/:
   profiles/launches_<tag>.csv (copy), profiles/ncu_<tag>_summary.csv (key metrics per captured kernel),
   profiles/traffic_<tag>.json (DRAM bytes per launch, read by bench.py for roofline.traffic), profiles/bench_<tag>.json."""
import csv, json, os, shutil, subprocess, sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)
for f in (f"launches_{tag}.csv", f"bench_{tag}.json", f"bench_ref_{tag}.json"):
    if os.path.exists(os.path.join(go, f)):
        shutil.copy(os.path.join(go, f), os.path.join(pr, f))

# ---- launch list: share of each kernel in the step
rows = list(csv.reader(l for l in open(os.path.join(go, f"launches_{tag}.csv")) if l.startswith('"')))
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    d[r[ki].split("(")[0].replace("void ", "").replace("vdbm::", "")].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
lines = ["kernel,launches,mean_us,total_us,share_of_gpu_time"]
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{k},{len(v)},{sum(v)/len(v)/1e3:.1f},{sum(v)/1e3:.1f},{sum(v)/tot:.3f}")
open(os.path.join(pr, f"launch_shares_{tag}.csv"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

# ---- full captures
import glob
hdr, units, data = None, None, []
for rep in sorted(glob.glob(os.path.join(go, f"prof_{tag}_*.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    if hdr is None:
        hdr, units = rows[0], rows[1]
    # different kernels expose the same metric set with --set full; align by name. ncu picks the UNIT of a metric per report
    # (the DDA's DRAM writes come in Kbyte, the update kernel's in Mbyte): rescale every value to the first report's unit.
    idxmap = {h: i for i, h in enumerate(rows[0])}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    def conv(val, u_from, u_to):
        if u_from == u_to or u_from not in scale or u_to not in scale:
            return val
        try:
            return "%.6f" % (float(val.replace(",", "")) * scale[u_from] / scale[u_to])
        except ValueError:
            return val
    for r in rows[2:]:
        data.append([conv(r[idxmap[h]], rows[1][idxmap[h]], units[j]) if h in idxmap else "" for j, h in enumerate(hdr)])
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_requests_srcunit_tex_op_red.sum", "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "sm__cycles_elapsed.max"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
out = [",".join(w for w, _ in idx), ",".join(units[i] for _, i in idx)]
traffic = {}
def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
for r in data:
    out.append(",".join('"%s"' % r[i].split("(")[0] if w == "Kernel Name" else r[i] for w, i in idx))
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("vdbm::", "")
    rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
    wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
    traffic.setdefault(name, []).append(rd + wr)
open(os.path.join(pr, f"ncu_{tag}_summary.csv"), "w").write("\n".join(out) + "\n")
tj = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)} for k, v in traffic.items()}
tj["_source"] = f"ncu --set full --clock-control none, prof_{tag}.ncu-rep (dram__bytes_read.sum + dram__bytes_write.sum), cfg2 steady-state scans"
json.dump(tj, open(os.path.join(pr, f"traffic_{tag}.json"), "w"), indent=1)
print(json.dumps(tj, indent=1))
