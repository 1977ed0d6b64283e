"""Diagnostic: key metrics + stall breakdown + hottest instructions of one kernel in an .ncu-rep."""
import csv, sys, subprocess, collections, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h, v = r[0], r[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in h:
        print(f"{k} = {v[h.index(k)]} {r[1][h.index(k)]}")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:])); hdr = rows[0]; rows = rows[1:]
cols = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
tot = collections.Counter()
for row in rows:
    for i in cols:
        tot[hdr[i]] += int(row[i] or 0)
s = sum(tot.values())
print("# warp-state samples:", "  ".join(f"{k[6:]} {c / s:.3f}" for k, c in tot.most_common(10)))
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
b = collections.Counter()
for row in rows:
    b[int(row[iE])] += 1
top = sorted(b.items(), key=lambda kv: -kv[0] * kv[1])[:4]
ti = sum(int(row[iE]) for row in rows)
print("# instruction buckets (exec count x static instrs, share):", "  ".join(f"{k}x{n} {k * n / ti:.3f}" for k, n in top))
if "--hot" in sys.argv:
    for row in sorted(rows, key=lambda q: -int(q[iN]))[:25]:
        st = sorted(((hdr[i][6:], int(row[i] or 0)) for i in cols if int(row[i] or 0) > 0), key=lambda kv: -kv[1])[:3]
        print(row[iN], row[iE], row[iS].strip()[:80], st)
