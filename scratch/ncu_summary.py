"""ncu -i REP --page raw --csv  ->  a small JSON with the metrics the profiles/ summaries quote (one list entry per launch)."""
import csv, io, json, subprocess, sys
rep, out, title, source = sys.argv[1:5]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__shared_mem_per_block_static", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
res = {"kernel": title, "source": source, "launches": [r[hdr.index("Kernel Name")][:60] for r in data], "metrics": {}}
for m in WANT:
    if m in hdr:
        i = hdr.index(m)
        res["metrics"][m + " [" + units[i] + "]"] = [float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None for r in data]
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res["metrics"], indent=1)[:3000])
