"""number of launches per step matching a regex in an ncu launch list that covers `steps` steps"""
import csv, io, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
n = sum(1 for row in csv.DictReader(io.StringIO("".join(lines))) if re.search(sys.argv[2], row["Kernel Name"]))
print(n // int(sys.argv[3]))
