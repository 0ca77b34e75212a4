"""DRAM traffic per C-ABI op from one `ncu --set full` capture of ONE train step (all library kernels, launch order):
    python tools/ncu_traffic.py gpurun_out/<x>.ncu-rep profiles/traffic_ppi.json
The kernels of a step run in the order  [proj_fwd, edge_fwd] x layers, then [edge_bwd, proj_bwd] x layers reversed;
consecutive launches of one class form one op.  bench.py reads the JSON to fill `roofline.traffic` (measured
dram__bytes_read.sum + dram__bytes_write.sum, summed over the op's kernels) next to the algorithmic bytes."""
import csv, io, json, os, subprocess, sys

# checked in this order ("gt_amax_kernel" must win over "amax_kernel")
CLASSES = {"EB": ("bwd_prep", "colsum_kernel", "edge_bwd_", "gt_amax_kernel", "bwd_finish_kernel"),
           "EF": ("edge_fwd_", "head_mean_kernel"),
           "P": ("amax_kernel", "split_kernel", "gemm_tc_kernel", "gemm_simt_kernel", "logits_")}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, out_path):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        return float(r[ix[key]].replace(",", "")) * SCALE.get(units[ix[key]], 1.0)
    runs = []   # [(class, [ (name, dram_bytes, us) ])]
    for r in body:
        name = r[ix["Kernel Name"]]
        cls = next((c for c, pats in CLASSES.items() if any(p in name for p in pats)), None)
        if cls is None:
            continue
        rec = (name.split("(")[0], val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
               float(r[ix["gpu__time_duration.sum"]].replace(",", "")) / (1e3 if units[ix["gpu__time_duration.sum"]] == "ns" else 1.0))
        if runs and runs[-1][0] == cls:
            runs[-1][1].append(rec)
        else:
            runs.append((cls, [rec]))
    layers = sum(1 for c, _ in runs if c == "EF")
    seq = [c for c, _ in runs]
    assert seq == ["P", "EF"] * layers + ["EB", "P"] * layers, f"unexpected launch order: {seq}"
    ops = {}
    for k, (cls, recs) in enumerate(runs):
        if k < 2 * layers:
            layer, op = k // 2, ("b200gat_proj_fwd" if cls == "P" else "b200gat_edge_fwd")
        else:
            j = k - 2 * layers
            layer, op = layers - 1 - j // 2, ("b200gat_edge_bwd" if cls == "EB" else "b200gat_proj_bwd")
        ops[f"{op}:{layer}"] = {"dram_bytes": sum(r[1] for r in recs), "us_under_ncu": sum(r[2] for r in recs),
                                "kernels": [r[0].replace("b200gat::", "").replace("void ", "") for r in recs]}
    sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or os.environ.get("B200GAT_GIT_SHA")
    json.dump({"source": rep, "git_sha": sha, "ops": ops}, open(out_path, "w"), indent=1)
    for k, v in ops.items():
        print(f"{k:24s} {v['dram_bytes'] / 1e6:9.1f} MB  {v['us_under_ncu']:8.1f} us  {len(v['kernels'])} kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
