"""Print the metrics that matter from an .ncu-rep (raw page) + top stall instructions (source page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1  # launch (row of the raw page) to summarise; source page = whole report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], (rows[2:] if len(rows) > 2 else rows[1:])[which]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__pcsamp_warps_issue_stalled"]
for h, v in zip(hdr, vals):
    if any(h.startswith(k) for k in keys) and "not_issued" not in h and ".per_second" not in h and "pct_of_peak_sustained_elapsed" not in h.replace("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "").replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","").replace("lts__throughput.avg.pct_of_peak_sustained_elapsed","").replace("l1tex__throughput.avg.pct_of_peak_sustained_elapsed",""):
        if h.startswith("smsp__pcsamp") and v in ("0", ""): continue
        print(f"{h} = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
if topn <= 0 or not data: sys.exit(0)
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print(f"--- source: {len(data)} SASS instructions, {tot} samples; top {topn} by samples")
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:topn]:
    print(f"{int(r[isamp] or 0):7d} {100*int(r[isamp] or 0)/max(tot,1):5.1f}%  exec {r[iex]:>10s}  {r[isrc][:90]}")
