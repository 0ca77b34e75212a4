"""Top warp-stall reasons and key throughput metrics per captured launch of an `ncu --set full` report.
usage: ncu_stalls.py <report.ncu-rep>"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
keys = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum"]
for r in body:
    print(r[ix["Kernel Name"]][:90])
    for k in keys:
        if k in ix:
            print(f"    {k:75s} {r[ix[k]]} {units[ix[k]]}")
    vals = sorted(((float(r[ix[k]].replace(",", "")), k) for k in stall if r[ix[k]] not in ("", "n/a")), reverse=True)
    for v, k in vals[:6]:
        print(f"    stall {v:7.2f}  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
