"""profiles/traffic_k_tc_bwd.json from an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,
sm__warps_active.avg.pct_of_peak_sustained_active) of tools/prof_one.py at the given point count; the file carries a hash of
siren_tc.cuh so that bench.py never quotes a capture of an older kernel.
usage: python tools/make_traffic_json.py <ncu.csv> <points> [out.json]"""
import csv, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, points = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic_k_tc_bwd.json")
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
h = rows[0]
ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[idi]), {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
pick = [d for i, d in sorted(launches.items()) if "k_tc_bwd" in d["name"] and d["name"].rstrip().endswith(", 0>(Params, float4 *)")]
if not pick:
    pick = [d for i, d in sorted(launches.items()) if "k_tc_bwd" in d["name"]]
k = pick[-1]                                    # last (warm) launch of the plain backward
sha = hashlib.sha256(open(os.path.join(ROOT, "insr_pde_b200", "csrc", "siren_tc.cuh"), "rb").read()).hexdigest()[:16]
rec = {"source": "ncu --clock-control none, tools/prof_one.py --points %d --reps 2 --lsq (round 2, final kernels)" % points,
       "points": points, "source_sha16": sha,
       "kernel": {"name": k["name"], "dram_read_bytes": k["dram__bytes_read.sum"] * (1e6 if k["dram__bytes_read.sum"] < 1e5 else 1),
                  "dram_write_bytes": k["dram__bytes_write.sum"] * (1e6 if k["dram__bytes_write.sum"] < 1e5 else 1),
                  "duration_ns": k.get("gpu__time_duration.sum"),
                  "pipe_tensor_pct": k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  "issue_pct": k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "warps_active_pct": k.get("sm__warps_active.avg.pct_of_peak_sustained_active")}}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec))
