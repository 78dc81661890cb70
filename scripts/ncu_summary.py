"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + stall breakdown + hottest SASS lines.
usage: ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
for r in rows[2:]:
    print("== launch")
    for h, u, v in zip(hdr, units, r):
        if h in KEYS:
            print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    def _isint(v):
        try:
            int(v or 0)
            return True
        except ValueError:
            return False
    hdr = rows[1]
    _c = hdr.index("Warp Stall Sampling (All Samples)")
    # several kernels in one report: every kernel's table repeats the header rows -- keep data rows only
    data = [r for r in rows[2:] if len(r) == len(hdr) and _isint(r[_c]) and r[0] != hdr[0]]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
    total = sum(int(r[idx["Warp Stall Sampling (All Samples)"]] or 0) for r in data)
    print("== warp stall samples (all):", total)
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {s:28s} {v:10d}  {100.0 * v / max(1, total):5.1f}%")
    print("== hottest SASS lines (samples, executed, instruction, top stall)")
    for r in sorted(data, key=lambda r: -int(r[idx["Warp Stall Sampling (All Samples)"]] or 0))[:15]:
        st = sorted(((s, int(r[idx[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[0]
        print(f"  {r[idx['Warp Stall Sampling (All Samples)']]:>8s} {r[idx['Instructions Executed']]:>10s}  {r[idx['Source']].strip()[:64]:64s} {st[0]}")
