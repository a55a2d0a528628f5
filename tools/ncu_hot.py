"""Hot spots of one kernel in an ncu report: top SASS instructions by stall samples + a coarse profile along the
instruction stream.  usage: ncu_hot.py report.ncu-rep kernel_regex [ntop]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(r for r in rows if "Source" in r and "# Samples" in r)
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for r in rows:
    if len(r) > max(isamp, iex) and r[isamp].isdigit():
        data.append(r)
    elif data and r and r[0] == "Kernel Name":
        break                                   # first matching launch only
tot = sum(int(r[isamp]) for r in data) or 1
totex = sum(int(r[iex]) for r in data) or 1
print("samples", tot, "sass instructions", len(data), "warp instructions executed", totex)
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[isamp]))[:ntop]:
    print(f"{int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}%  #{idx[id(r)]:5d} ex={r[iex]:>8}  {r[isrc].strip()[:110]}")
print()
step = max(100, len(data) // 40)
for b in range(0, len(data), step):
    s = sum(int(r[isamp]) for r in data[b:b + step]); e = sum(int(r[iex]) for r in data[b:b + step])
    print(f"sass {b:5d}-{b+step:5d}: samples {100*s/tot:5.1f}%  executed {100*e/totex:5.1f}%   first: {data[b][isrc].strip()[:60]}")
