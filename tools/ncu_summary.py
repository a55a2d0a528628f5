"""Summarise an ncu --set full report (read on the CPU box): per kernel duration, pipe utilisation,
shared-memory wavefronts, DRAM bytes, registers and the top warp-stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg']
stall = [n for n in h if n.startswith('smsp__pcsamp_warps_issue_stalled') and not n.endswith('_not_issued')]
for r in rows[2:]:
    print("==", r[h.index('Kernel Name')], "  [units row:", rows[1][h.index('gpu__time_duration.sum')], "]")
    for w in want:
        if w in h:
            print(f"   {w:80s} {r[h.index(w)]} {rows[1][h.index(w)]}")
    vals = []
    for n in stall:
        try:
            vals.append((float(r[h.index(n)]), n.replace('smsp__pcsamp_warps_issue_stalled_', '')))
        except ValueError:
            pass
    tot = sum(v for v, _ in vals) or 1
    print("   stalls: " + ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in sorted(vals, reverse=True)[:7]))
