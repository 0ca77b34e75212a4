"""bench.py — the reference's headline metric on the B200-native path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ppi|cifar|cora|heads|large]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): GAT layer fwd+bwd edges/s — unit of work = one processed edge of E' = E + N (self loops are
real work) per layer; a "step" is one full train step (zero_grad, forward, loss, backward, gradient all-reduce when
N > 1, Adam step — run_inductive.py:75-85) of the workload's model over one synthetic batch;
value = n_gpus * sum_layers E' / t_step.  Default workload = BASELINE.json configs[1]: the PPI-shaped inductive batch
(24 graphs, 56,944 nodes, 818,716 edges, 50 feats, 121 labels), 3-layer GAT 4/4/6 heads x 256.

One JSON line on stdout (rank 0).  `value` times the step with inputs resident in HBM and the CSR cached (the same
edge_index object every step, as in run_inductive.py:77); `e2e` times the same step from PINNED HOST buffers: H2D of
x / edge_index / y, CSR build, train step, D2H of the loss — every step.  `roofline` is the dominant ABI op of the step,
timed live with CUDA events around each C-ABI call; `kernels` lists every op.  `cpu_baseline` / `--impl reference` time
the CPU oracle port of the reference (oracle/gat_port.py; the reference itself needs torch_geometric, which is not
installable here, and /root/reference does not exist on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from atmlgraphattentionnetworks_b200 import synth  # noqa: E402

METRIC = "gat_layer_fwd_bwd_edges_per_s"
UNIT = "edges/s"


# ------------------------------------------------------------------------------------------------ workloads
HEADS = 8   # --heads: the point of the heads sweep (BASELINE configs[3]: 1, 2, 4, 8, 16 heads x 64)


def make_workload(name, seed, sample=None):
    """-> (data, spec, loss_fn, description).  `sample` bounds the CPU baseline (fraction of the batch)."""
    import torch.nn.functional as F
    if name == "ppi":
        data = synth.ppi_shaped(seed=seed, keep_graphs=sample)
        spec = synth.PPI_STACK
        loss_fn = lambda out, y: F.binary_cross_entropy_with_logits(out, y)   # noqa: E731
        desc = "PPI-shaped inductive batch (24 graphs, 56944 nodes, 818716 edges, 50 feats, 121 labels), 3-layer GAT 4/4/6 heads x 256/256/121"
    elif name == "large":
        data = synth.powerlaw(seed=seed) if sample is None else synth.powerlaw(seed=seed, num_nodes=75_000, num_edges=1_937_500)
        spec = synth.LARGE_STACK
        loss_fn = lambda out, y: F.nll_loss(F.log_softmax(out, dim=1), y)     # noqa: E731
        desc = "large power-law graph (2.4M nodes, 62M edges, 100 feats, 47 classes), 3-layer GAT 4 heads x 128"
    elif name == "heads":
        data = synth.ppi_shaped(seed=seed, keep_graphs=sample)
        spec = [(50, 64, HEADS, True)]
        loss_fn = lambda out, y: out.sum()                                     # noqa: E731
        desc = f"heads sweep point: one layer 50 -> {HEADS} heads x 64 on the PPI-shaped batch"
    elif name == "cifar":
        # BASELINE configs[2]: run_gnn_benchmark.py:35-66 — GATNet('GAT','CIFAR10',F): conv1 -> elu -> conv2 -> elu ->
        # per-graph mean -> lin1 -> relu -> lin2 -> log_softmax, nll_loss per graph; dropout 0.0 (GATNet.py:19-20)
        data = synth.cifar_shaped(seed=seed, num_graphs=128 if sample is None else 16)
        spec = [(5, 8, 8, True), (64, 8, 8, True)]
        loss_fn = lambda out, y: F.nll_loss(out, y)                            # noqa: E731
        desc = "CIFAR10-superpixel-shaped batch (128 graphs x ~117 nodes, kNN k=8, 5 feats, 10 classes), GATNet('GAT','CIFAR10',5)"
    elif name == "cora":
        # BASELINE configs[0]: run_inductive.py:75-85 — GATNet('GAT','Cora',1433) in training mode (feature dropout 0.6 and
        # attention dropout 0.6 are active, as in the reference's train step), nll_loss over all nodes
        data = synth.cora_shaped(seed=seed)
        spec = [(1433, 8, 8, True), (64, 7, 1, False)]
        loss_fn = lambda out, y: F.nll_loss(out, y)                            # noqa: E731
        desc = "Cora-shaped graph (2708 nodes, 10556 edges, 1433 feats, 7 classes), GATNet('GAT','Cora',1433), dropout 0.6 active"
    else:
        raise ValueError(name)
    return data, spec, loss_fn, desc


def layer_edges(data):
    return int(data.edge_index.shape[1] + data.x.shape[0])


# ------------------------------------------------------------------------------------ algorithmic bytes
def algorithmic_bytes(op, n, ep, f, c, h, concat, need_gx, cached):
    """Bytes one ABI op must move (fp32 values, int32 indices); G = gather multiplicity: N when the gathered operand
    of the largest graph block fits half of L2 ("cached regime": the block-diagonal PPI-shaped batch, blocks <= 15 MB),
    E' otherwise ("streaming regime": the 2.4M-node graph, 4.9 GB of Wh) — SURVEY.md §8d."""
    d = h * c
    d_out = d if concat else c
    g = n if cached else ep
    if op == "b200gat_proj_fwd":
        return 4 * (n * f + d * f + n * d + 2 * n * h), "tensor"
    if op == "b200gat_edge_fwd":
        extra = 4 * n * d if (not concat and h > 1) else 0
        return 4 * ((n + 1) + ep + 2 * n * h + n * d_out) + 4 * g * (d + h) + extra, "hbm"
    if op == "b200gat_edge_bwd":
        return 4 * ((n + 1) + ep + n * d_out + 3 * n * d + 4 * n * h) + 4 * g * (d_out + 4 * h), "hbm"
    if op == "b200gat_proj_bwd":
        return 4 * (n * d + (2 if need_gx else 1) * n * f + 2 * d * f), "tensor"
    raise KeyError(op)


def gemm_flops(op, n, f, c, h, need_gx):
    d = h * c
    if op == "b200gat_proj_fwd":
        return 2.0 * n * f * d + 4.0 * n * d
    if op == "b200gat_proj_bwd":
        return 2.0 * n * f * d * (2 if need_gx else 1)
    return 0.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sustained=float(p["bf16_tflops_sustained"]),
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as e:   # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:   # pragma: no cover
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_reference_leg(workload, steps, warmup, sample):
    """The reference's CPU PyTorch path (oracle port of GAT.py, same op sequence, autograd backward) on all host
    cores, on a bounded sample of the workload.  -> (edges/s, seconds per step, description)."""
    from oracle.gat_port import PortGATNet, PortStack
    torch.set_num_threads(os.cpu_count() or 1)
    data, spec, loss_fn, _ = make_workload(workload, 0, sample=sample)
    torch.manual_seed(0)
    if workload in ("cifar", "cora"):
        net = PortGATNet("GAT", "CIFAR10" if workload == "cifar" else "Cora", data.x.shape[1]).train()
        model = lambda x, ei: net(data)                                        # noqa: E731
        opt = torch.optim.Adam(net.parameters(), lr=5e-3, weight_decay=5e-4)
    else:
        model = PortStack(spec, dropout=0.0)
        opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=5e-4)
    ep = layer_edges(data)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(data.x, data.edge_index), data.y)
        loss.backward()
        opt.step()
        return float(loss)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    what = (f"{data.num_graphs} of 24 graphs of the PPI-shaped batch" if workload in ("ppi", "heads") else
            f"{data.num_graphs} of 128 graphs of the CIFAR-shaped batch" if workload == "cifar" else
            "the whole Cora-shaped graph" if workload == "cora" else
            "1/32-scale graph from the same power-law generator")
    desc = f"{what}: {data.x.shape[0]} nodes, {ep} edges incl. self loops, {len(spec)} layers, {steps} steps after {warmup} warm-up"
    return len(spec) * ep / dt, dt, desc


# ------------------------------------------------------------------------------------------------ main
def _emit(line):
    """The ONE JSON line goes to the real stdout; everything else that lands on fd 1 while the bench runs (NCCL prints
    its version banner there) is diverted to stderr so that the line stays machine-readable."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ppi", choices=["ppi", "heads", "large", "cifar", "cora"])
    ap.add_argument("--heads", type=int, default=8, help="--workload heads: number of heads (x 64 channels)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="capture the resident train step in a CUDA graph and time replays (launch-bound workloads)")
    ap.add_argument("--cpu-sample-graphs", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="resident leg only (for ncu runs): warm-up + steps, minimal JSON")
    args = ap.parse_args()
    global HEADS
    HEADS = args.heads
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = max(args.cpu_sample_graphs, 4) if args.workload != "large" else 1
        val, dt, desc = cpu_reference_leg(args.workload, args.steps, max(args.warmup, 1), sample)
        _, _, _, wdesc = make_workload(args.workload, 0, sample=sample)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": max(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wdesc, "cpu_sample": desc},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": desc},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the GAT hot path has no CPU fallback); use --impl reference for the CPU leg")
    import torch.distributed as dist
    from atmlgraphattentionnetworks_b200 import _abi
    from atmlgraphattentionnetworks_b200.gatnet import GATNet, GATStack
    from atmlgraphattentionnetworks_b200.graph import GLOBAL_CACHE
    from atmlgraphattentionnetworks_b200.parallel import GradBucket, all_reduce_packed_grads

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus} (launch with torchrun)"

    # graph batches: weak scaling — every rank owns its own batch of independent graphs (seed = rank), model replicated.
    # one large graph (--workload large, N > 1): strong scaling — destination-row partition, Wh all-gather per layer.
    partitioned = args.workload == "large" and world > 1
    data, spec, loss_fn, wdesc = make_workload(args.workload, seed=0 if partitioned else rank)
    ep = layer_edges(data)
    n = data.x.shape[0]
    torch.manual_seed(0)
    is_net = args.workload in ("cifar", "cora")
    if is_net:
        from types import SimpleNamespace
        net = GATNet("GAT", "CIFAR10" if args.workload == "cifar" else "Cora", data.x.shape[1]).to(dev).train()
        batch_d = data.batch.to(dev) if hasattr(data, "batch") else None
        model = lambda x, ei: net(SimpleNamespace(x=x, edge_index=ei, batch=batch_d, num_graphs=data.num_graphs))   # noqa: E731
        model.parameters = net.parameters
    else:
        model = GATStack(spec, dropout=0.0).to(dev)
    # gradients are set to None every step and autograd adopts the kernels' packed output buffers as .grad without a copy;
    # N > 1 (graph batches): those few buffers are all-reduced in ONE NCCL group (parallel.all_reduce_packed_grads).
    # The row-partitioned large graph keeps the flat bucket (its stage functions build gradients through torch ops).
    bucket = GradBucket(model.parameters()) if (world > 1 and args.workload == "large") else None
    params = list(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=5e-4, fused=True,   # run_inductive.py:18-19,65
                           capturable=args.cuda_graph)
    part = None
    if partitioned:
        from atmlgraphattentionnetworks_b200.partition import PartitionedGATStack, build_row_partition
        part = build_row_partition(data.edge_index.to(dev), n, world, rank)
        pmodel = PartitionedGATStack(model)
        x_h, y_h = data.x[part.lo:part.hi].contiguous().pin_memory(), data.y[part.lo:part.hi].contiguous().pin_memory()
        ei_h = torch.zeros((2, 0), dtype=torch.int64).pin_memory()      # the partitioned graph is static and resident
        torch.cuda.empty_cache()
    else:
        x_h, ei_h, y_h = data.x.pin_memory(), data.edge_index.pin_memory(), data.y.pin_memory()
    x_d, ei_d, y_d = x_h.to(dev), ei_h.to(dev), y_h.to(dev)

    def train_step(x, ei, y):
        if bucket is not None:
            bucket.zero()
        else:
            opt.zero_grad(set_to_none=True)
        if partitioned:
            out = pmodel(x, part)
            # global mean loss = sum over ranks of (own sum / N); parameter gradients are then SUMMED over ranks
            loss = torch.nn.functional.nll_loss(torch.nn.functional.log_softmax(out, dim=1), y, reduction="sum") / n
            loss.backward()
            bucket.all_reduce_mean(weight=1.0)
        else:
            out = model(x, ei)
            loss = loss_fn(out, y)
            loss.backward()
            if world > 1:
                all_reduce_packed_grads(params)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    warmup = max(args.warmup, 3)
    # ---- resident leg ----
    for _ in range(warmup):
        train_step(x_d, ei_d, y_d)
    resident = lambda: train_step(x_d, ei_d, y_d)                               # noqa: E731
    launches_per_replay = None
    if args.cuda_graph:
        # whole-step capture: zero_grad + forward + loss + backward + fused Adam of the resident
        # batch become ONE graph launch; the CSR is cached (no host sync inside), every buffer comes from the graph's pool
        assert world == 1, "--cuda-graph is a single-GPU mode"
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                train_step(x_d, ei_d, y_d)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        l0 = _abi.launch_count()
        with torch.cuda.graph(graph):
            static_loss = train_step(x_d, ei_d, y_d)
        launches_per_replay = _abi.launch_count() - l0
        resident = graph.replay
        for _ in range(3):
            resident()
    launches0 = _abi.launch_count()
    with ClockSampler(local_rank) as clocks:
        ms_step = timed(resident, args.steps)
    launches = _abi.launch_count() - launches0
    if launches_per_replay is not None:
        launches = launches_per_replay * args.steps
    if args.profile:
        if rank == 0:
            _emit({"profile_run": True, "ms_per_step": ms_step, "gpu_launches": launches})
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- end-to-end leg: pinned host buffers in, loss out, every step (new tensors => CSR rebuilt every step) ----
    # The host->device copy of step k+1 is issued on a copy stream while step k computes (what a data loader with a
    # prefetch depth of one does): every step still copies its own inputs from pinned host memory inside the timed
    # region, builds its CSR from the fresh edge_index, and reads its loss back.
    copy_stream = torch.cuda.Stream()
    # two static sets of device buffers (no per-step allocation on the copy stream: cross-stream frees make the caching
    # allocator synchronise, which cost 4 ms per step next to NCCL); the in-place copy bumps edge_index's version
    # counter, so the graph cache misses and the CSR is rebuilt for every batch
    slots = [(torch.empty_like(x_d), torch.empty_like(ei_d), torch.empty_like(y_d)) for _ in range(2)]
    consumed = [None, None]          # event: the step that read slot k has finished

    loss_h = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

    def upload(k):
        """batch k: pinned host -> device slot k % 2, then graph ingestion (GAT.py:38 + the CSR / CSC / degree-class build:
        b200gat_csr_build, b200gat_hub_rows, one status read) — all on the copy stream, while step k - 1 computes."""
        x, ei, y = slots[k % 2]
        with torch.cuda.stream(copy_stream):
            if consumed[k % 2] is not None:
                copy_stream.wait_event(consumed[k % 2])
            x.copy_(x_h, non_blocking=True)
            ei.copy_(ei_h, non_blocking=True)
            y.copy_(y_h, non_blocking=True)
            if not partitioned:
                GLOBAL_CACHE.get(ei, n)       # the layers of step k find this batch's CSR in the cache
            done = torch.cuda.Event()
            done.record(copy_stream)
        return (x, ei, y), done

    def e2e_run(steps):
        nxt = upload(0)
        pending = None                       # (event, pinned scalar) of the previous step's loss
        losses = []
        for k in range(steps):
            (x, ei, y), done = nxt
            torch.cuda.current_stream().wait_event(done)
            loss = train_step(x, ei, y)
            loss_h[k % 2].copy_(loss.detach(), non_blocking=True)      # device -> pinned host, read one step later
            ev = torch.cuda.Event()
            ev.record()
            consumed[k % 2] = ev
            if k + 1 < steps:
                nxt = upload(k + 1)          # overlaps step k on the GPU (the host blocks on the ingestion status read)
            if pending is not None:
                pending[0].synchronize()
                losses.append(float(pending[1]))
            pending = (ev, loss_h[k % 2])
        pending[0].synchronize()
        losses.append(float(pending[1]))
        assert len(losses) == steps and all(v == v for v in losses)
    e2e_run(4)                           # warm-up: also fills the copy stream's allocator pool (CSR arrays, sort workspace)
    e2e_steps = max(args.steps // 2, 3)
    # median of three timed repetitions (host-side jitter of the per-step synchronisations is +-1 ms run to run)
    ms_e2e = statistics.median(timed(lambda: e2e_run(e2e_steps), 1) / e2e_steps for _ in range(3))
    h2d = x_h.numel() * 4 + ei_h.numel() * 8 + y_h.numel() * y_h.element_size()

    # ---- per-op breakdown (CUDA events around each C-ABI call) ----
    GLOBAL_CACHE.clear()
    train_step(x_d, ei_d, y_d)
    _abi.timing = []
    reps = 5
    for _ in range(reps):
        train_step(x_d, ei_d, y_d)
    torch.cuda.synchronize()
    per_raw = {}
    for name, geom, s, e in _abi.timing:
        per_raw.setdefault((name, geom), []).append(s.elapsed_time(e))
    _abi.timing = None
    per = {}
    for (name, geom), ts in per_raw.items():      # staged backward (partitioned mode): prep + csc + finish = edge_bwd
        key = ("b200gat_edge_bwd" if name.startswith("b200gat_edge_bwd") else name, geom)
        per[key] = [a + b for a, b in zip(per[key], ts)] if key in per else list(ts)
    n_loc = part.n_own if partitioned else n
    ep_loc = int(part.col.numel() + part.crow.numel()) // 2 if partitioned else ep
    peaks = measured_peaks()
    kernels = []
    for li, (f, c, h, concat) in enumerate(spec):
        for op in ("b200gat_proj_fwd", "b200gat_edge_fwd", "b200gat_edge_bwd", "b200gat_proj_bwd"):
            ts = per.get((op, (f, c, h, bool(concat))), [])
            if not ts:
                continue
            ms = statistics.median(ts)
            need_gx = li > 0
            nbytes, bound = algorithmic_bytes(op, n_loc, ep_loc, f, c, h, concat, need_gx, cached=args.workload != "large")
            rec = {"op": op, "layer": li, "geom": f"{f}->{h}x{c}{'cat' if concat else 'mean'}", "ms": ms,
                   "alg_bytes": nbytes, "GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peaks["hbm"]}
            fl = gemm_flops(op, n_loc, f, c, h, need_gx)
            if fl:
                rec["TFLOPs"] = fl / ms / 1e9
            kernels.append(rec)
    dom = max(kernels, key=lambda r: r["ms"]) if kernels else None
    # measured DRAM traffic of each op (one committed `ncu --set full` capture of this workload, tools/ncu_traffic.py)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.isfile(tpath) and not partitioned:
        traffic = json.load(open(tpath)).get("ops", {})
    for rec in kernels:
        t = traffic.get(f"{rec['op']}:{rec['layer']}")
        rec["dram_traffic_bytes"] = t["dram_bytes"] if t else None
    roofline = None
    if dom is not None:
        if dom["op"].startswith("b200gat_edge"):
            roofline = {"bound": "hbm", "achieved": dom["GBps"], "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": dom["GBps"] / peaks["hbm"], "traffic": dom["dram_traffic_bytes"],
                        "alg_bytes": dom["alg_bytes"]}
        else:
            roofline = {"bound": "tensor", "achieved": dom.get("TFLOPs", 0.0), "peak": peaks["bf16_sustained"],
                        "unit": "TFLOP/s", "frac": dom.get("TFLOPs", 0.0) / peaks["bf16_sustained"],
                        "traffic": dom["dram_traffic_bytes"], "alg_bytes": dom["alg_bytes"]}
        roofline.update({"kernel": f"{dom['op']} layer {dom['layer']} ({dom['geom']})", "ms": dom["ms"],
                         "peak_source": peaks["source"] + " (of measured)"})
        if args.workload != "large" and roofline["bound"] == "hbm":
            # SURVEY.md §8d caveat 2: in the cached regime the HBM count is the compulsory traffic; every edge still pulls
            # its row through L2 -> SM (4.4x the HBM bytes on the PPI-shaped layers), which is what bounds these kernels
            roofline["regime"] = "cached: gathered rows are L2-resident, the kernel is bound by L2->SM gather traffic (DESIGN.md §5)"
        elif roofline["bound"] == "hbm":
            roofline["regime"] = "streaming: every gathered row comes from HBM"
    edge_ms = sum(r["ms"] for r in kernels if r["op"].startswith("b200gat_edge"))
    edge_bytes = sum(r["alg_bytes"] for r in kernels if r["op"].startswith("b200gat_edge"))
    edge_phase = {"per_rank": partitioned, "ms": edge_ms, "alg_bytes": edge_bytes, "GBps": edge_bytes / edge_ms / 1e6 if edge_ms else None,
                  "frac_of_measured_hbm": edge_bytes / edge_ms / 1e6 / peaks["hbm"] if edge_ms else None,
                  "frac_of_nominal_8TBps": edge_bytes / edge_ms / 1e6 / 8000.0 if edge_ms else None,
                  "edges_per_s": len(spec) * ep_loc / (edge_ms / 1e3) if edge_ms else None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N = 1 only
        val, dt, desc = cpu_reference_leg(args.workload, 3, 1, args.cpu_sample_graphs if args.workload != "large" else 1)
        cpu = {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": desc,
               "s_per_step": dt}
    total_edges = len(spec) * ep * (1 if partitioned else world)
    line = {
        "metric": METRIC, "value": total_edges / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if partitioned else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wdesc,
                   "parallelism": (f"row{world} (destination-row partition, NCCL all-gather of Wh / gout per layer, reduce-scatter of g_s_dst, grad all-reduce)"
                                   if partitioned else f"dp{world} (independent graph batches per rank, one-group NCCL all-reduce of the packed gradient buffers)"),
                   "edges_per_layer_incl_self_loops": ep, "input_edges": int(data.edge_index.shape[1]), "nodes": n,
                   "layers": len(spec), "step": "zero_grad + fwd + loss + bwd + grad all-reduce + fused Adam",
                   "cuda_graph": bool(args.cuda_graph),
                   "l2": "inputs larger than L2 (per-step working set ~2 GB vs 126 MB L2); no explicit flush"},
        "e2e": {"value": total_edges / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "includes": "every step: H2D of its batch from pinned host and graph ingestion (CSR/CSC build, 2 radix sorts, degree classes) on a copy stream one batch ahead, train step, loss D2H into pinned memory read one step behind"},
        "gpu_launches": launches, "roofline": roofline, "edge_phase": edge_phase, "kernels": kernels,
        "cpu_baseline": cpu, "clocks": clocks.summary(),
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
