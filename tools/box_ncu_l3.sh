#!/bin/bash
# ncu --set full of the mean-layer (4 x 47) kernels of the 2.4 M-node graph's last layer, one step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"edge_bwd_mean_kernel|edge_fwd_row_stream_kernel|bwd_prep_mean_rows" -s 9 -c 3 -f -o /tmp/l3_prof \
    python bench.py --profile --steps 1 --warmup 3 --workload large > gpurun_out/r2y_ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_stalls.py /tmp/l3_prof.ncu-rep > gpurun_out/r2y_stalls.txt
python tools/ncu_summary.py /tmp/l3_prof.ncu-rep > gpurun_out/r2y_summary.md
ncu -i /tmp/l3_prof.ncu-rep --page source --csv > /tmp/l3_src.csv 2>/dev/null
python - <<'P'
import csv
rows = [r for r in csv.reader(open("/tmp/l3_src.csv"))]
out = open("gpurun_out/r2y_hot.txt", "w")
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]; hdr = rows[k + 1]
        ie, isrc, iss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
        body = []
        k += 2
        while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
            if rows[k] and rows[k][0].startswith("0x"): body.append(rows[k])
            k += 1
        tot = sum(int(r[iss]) for r in body) or 1
        out.write(f"== {name[:100]}  samples {tot}\n")
        for r in sorted(body, key=lambda r: -int(r[iss]))[:22]:
            out.write(f"   {int(r[iss]) / tot * 100:5.1f}%  x{r[ie]:>9s}  {r[isrc].strip()}\n")
    else:
        k += 1
P
cat gpurun_out/r2y_summary.md | tail -5; grep -A22 "^void\|stall" gpurun_out/r2y_stalls.txt | head -80
