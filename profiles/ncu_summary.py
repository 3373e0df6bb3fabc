#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, opcode mix, stall mix and
the hottest source lines.  Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [n_warps]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print("== headline metrics ==")
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
for k in want:
    if k in d:
        print(f"{k} [{d[k][0]}] = {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
op, tot = collections.Counter(), 0
stall = collections.Counter()
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    if len(r) <= ie:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
    if not m:
        continue
    n = int(r[ie] or 0)
    op[m.group(2).split(".")[0]] += n
    tot += n
    for h in names:
        stall[h] += int(r[hdr.index(h)] or 0)
nw = nwarps or 1.0
print(f"\n== opcode mix (warp instructions{' per warp' if nwarps else ''}; total {tot / nw:.0f}) ==")
for o, n in op.most_common(16):
    print(f"{o:8s} {n / nw:12.0f} {100 * n / tot:5.1f}%")
s = sum(stall.values())
print("\n== warp stall samples ==")
for h, n in stall.most_common(8):
    print(f"{h:26s} {100 * n / s:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, agg, tot, tots = None, {}, 0, 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] == "Line No" or r[2] != "-":
        continue
    try:
        ln, n, sm = int(r[0]), int(r[7] or 0), int(r[6] or 0)
    except ValueError:
        continue
    agg[(cur, ln)] = (n, sm, r[1].strip()[:80])
    tot += n
    tots += sm
print("\n== instructions by file ==")
ft = collections.Counter()
for (f, l), (n, sm, _) in agg.items():
    ft[f] += n
for f, n in ft.most_common():
    print(f"{f:18s} {n / nw:10.0f} {100 * n / tot:5.1f}%")
print("\n== hottest source lines (by stall samples) ==")
for (f, l), (n, sm, text) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{f}:{l:<4d} {n / nw:8.0f} {100 * n / tot:5.1f}% smp {100 * sm / max(tots, 1):5.1f}%  {text}")
print("\n== source lines by instructions executed ==")
for (f, l), (n, sm, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{f}:{l:<4d} {n / nw:8.0f} {100 * n / tot:5.1f}% smp {100 * sm / max(tots, 1):5.1f}%  {text}")
