#!/bin/bash
# Runs ON the GPU box: ncu --set full of the projection GEMMs of one step; summarised here (the report is too big to return)
# usage: box_ncu_gemm.sh <tag> <workload> <skip> <count>
TAG=$1; WL=$2; SKIP=$3; COUNT=$4
mkdir -p gpurun_out
python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $SKIP -c $COUNT -f -o /tmp/${TAG}_prof \
    python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_stalls.py /tmp/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_stalls.txt
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv > /tmp/${TAG}_src.csv 2>/dev/null
python - <<P
import csv
rows = [r for r in csv.reader(open("/tmp/${TAG}_src.csv"))]
out = open("gpurun_out/${TAG}_hot.txt", "w")
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
        for r in sorted(body, key=lambda r: -int(r[iss]))[:25]:
            out.write(f"   {int(r[iss]) / tot * 100:5.1f}%  x{r[ie]:>9s}  {r[isrc].strip()}\n")
    else:
        k += 1
P
