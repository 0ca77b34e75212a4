"""Summarise an `ncu --set full` report (read here, no GPU needed): one line per captured launch with the metrics the
roofline discussion uses.   python tools/ncu_summary.py gpurun_out/<x>.ncu-rep > profiles/<x>_summary.md"""
import csv, io, subprocess, sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "dram_rd_MB", None), ("dram__bytes_write.sum", "dram_wr_MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1), ("lts__t_bytes.sum", "l2_MB", None),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%", 1), ("l1tex__t_sector_hit_rate.pct", "l1_hit_%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1)]
UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{path}` ({len(body)} launches; cold-cache, serialised replays)\n")
    names = ["kernel"] + [c[1] for c in COLS if c[0] in idx]
    print("| " + " | ".join(names) + " |")
    print("|" + "---|" * len(names))
    for r in body:
        name = r[idx["Kernel Name"]]
        name = name.replace("b200gat::", "").replace("void ", "")
        name = name.split("(")[0][:60]
        out = [name]
        for key, label, _ in COLS:
            if key not in idx:
                continue
            v = r[idx[key]].replace(",", "")
            u = units[idx[key]]
            try:
                f = float(v)
            except ValueError:
                out.append(v); continue
            if label.endswith("_MB"):
                f *= UNIT_SCALE.get(u, 1.0)
                out.append(f"{f:.1f}")
            elif label == "us":
                f *= UNIT_SCALE.get(u, 1.0)
                out.append(f"{f:.1f}")
            elif label in ("regs", "grid", "block"):
                out.append(f"{int(f)}")
            else:
                out.append(f"{f:.1f}")
        print("| " + " | ".join(out) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
