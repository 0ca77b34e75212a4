"""SASS evidence table: python tools/sass_summary.py [libb200gat.so] > profiles/<round>_sass_summary.md
Mnemonic counts per kernel from `cuobjdump -sass` (sm_100a): tcgen05.mma = UTCHMMA, TMA loads = UTMALDG, tcgen05.ld = LDTM,
tcgen05.commit = UTCBAR, 128-bit global loads, red.global.add, MUFU.EX2, spills (STL / LDL)."""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "atmlgraphattentionnetworks_b200/libb200gat.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c["n"] += 1
        for key, pat in (("mma", r"^UTCHMMA"), ("mma2", r"^UTCHMMA.*2CTA"), ("tma", r"^UTMALDG"), ("ldtm", r"^LDTM"),
                         ("commit", r"^UTCBAR"), ("ldg128", r"^LDG\.E.*\.128"), ("stg128", r"^STG\.E.*\.128"),
                         ("red", r"^(RED|ATOMG)"), ("ex2", r"^MUFU\.EX2"), ("spill", r"^(STL|LDL)"), ("shfl", r"^SHFL")):
            if re.search(pat, op):
                c[key] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
total = collections.Counter()
for c in counts.values():
    total.update(c)
print(f"# SASS evidence: `cuobjdump -sass {lib}` (sm_100a), mnemonic counts per kernel (tools/sass_summary.py)\n")
print("Whole library: " + ", ".join(f"{k} x{total[k]}" for k in ("mma", "mma2", "tma", "ldtm", "commit", "ldg128", "stg128", "red", "ex2", "spill")) +
      "  (mma = UTCHMMA = tcgen05.mma, mma2 = its .2CTA form, tma = UTMALDG, ldtm = LDTM = tcgen05.ld, commit = UTCBAR, "
      "red = RED / ATOMG, spill = STL / LDL)\n")
print("| kernel | SASS instr | UTCHMMA | .2CTA | UTMALDG | LDTM | UTCBAR | LDG.128 | STG.128 | RED/ATOM | MUFU.EX2 | SHFL | STL/LDL |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for (mang, c), d in zip(counts.items(), dem):
    if "cub::" in d or "thrust::" in d:
        continue                                            # CUB's radix sort (csr_build) is library code
    d = re.sub(r"\((int|bool|unsigned int)\)", "", d)
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("b200gat::", "")
    print(f"| `{short}` | {c['n']} | {c['mma']} | {c['mma2']} | {c['tma']} | {c['ldtm']} | {c['commit']} | {c['ldg128']} | {c['stg128']} | "
          f"{c['red']} | {c['ex2']} | {c['shfl']} | {c['spill']} |")
