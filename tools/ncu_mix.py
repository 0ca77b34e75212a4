"""Instruction mix and hot instructions of one kernel from `ncu -i rep --page source --csv` output.
usage: ncu_mix.py <source.csv> [top_n]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1]))]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
ie, isrc, iss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie]) for r in body)
samp = sum(int(r[iss]) for r in body)
print("kernel", rows[0][1][:100] if rows[0] else "", "| warp instructions", tot, "| stall samples", samp)
mix = collections.Counter(); smix = collections.Counter()
for r in body:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op.split(".")[0]] += int(r[ie]); smix[op.split(".")[0]] += int(r[iss])
for k, v in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{k:10s} {v / tot * 100:5.1f}% of instr   {smix[k] / max(samp, 1) * 100:5.1f}% of stall samples")
