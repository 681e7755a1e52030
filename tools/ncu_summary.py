"""Text summary of an ncu report: headline metrics per launch + top stall sites (needs `ncu` on PATH).
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}")
for r in rows[2:]:
    print(f"\n== launch: {r[ci['Kernel Name']][:110]}")
    for k in KEYS:
        if k in ci:
            print(f"  {k:74s} {r[ci[k]]:>16s} {units[ci[k]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:]:
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    if len(lines) < 3:
        continue
    name = lines[0][1] if len(lines[0]) > 1 else "?"
    h = lines[1]
    c = {x: i for i, x in enumerate(h)}
    if "# Samples" not in c:
        continue
    data = [r for r in lines[2:] if len(r) == len(h)]
    stalls = [x for x in h if x.startswith("stall_") and "Not" not in x]
    tot = sum(int(r[c["# Samples"]] or 0) for r in data)
    agg = {s: sum(int(r[c[s]] or 0) for r in data) for s in stalls}
    print(f"\n== warp-state samples: {name[:100]}\n  total {tot}; " + ", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 50 > tot))
    print("  top stall sites (samples, executions, SASS):")
    for r in sorted(data, key=lambda r: -int(r[c["# Samples"]] or 0))[:12]:
        print(f"    {r[c['# Samples']]:>6s} {r[c['Instructions Executed']]:>10s}  {r[c['Source']].strip()[:90]}")
