"""bench.py — the reference's headline metric on the B200-native path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload all|ppi|large|cifar|cora|heads]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): GAT layer fwd+bwd edges/s — unit of work = one processed edge of E' = E + N (self loops are
real work) per layer; a "step" is one full train step (zero_grad, forward, loss, backward, gradient all-reduce when
N > 1, Adam step — run_inductive.py:75-85) of the workload's model over one synthetic batch;
value = sum over ranks and layers of E' / t_step.

ONE JSON line on stdout (rank 0).  Its top-level keys are the HEADLINE workload = BASELINE.json configs[1]: the PPI-shaped
inductive batch (24 graphs, 56,944 nodes, 818,716 edges, 50 feats, 121 labels), 3-layer GAT 4/4/6 heads x 256 (`value` =
inputs resident in HBM, CSR cached; `e2e` = the same step from PINNED HOST buffers: H2D of x / edge_index / y, CSR build,
train step, D2H of the loss, every step).  With the default `--workload all` the same line also carries one sub-record per
other BASELINE config, each measured the same way in the same process:
    "large"  configs[4]  2.4 M-node power-law graph, 3-layer GAT 4 x 128: 1 GPU resident; row-partitioned at N > 1 (strong)
    "cifar"  configs[2]  CIFAR10-superpixel-shaped batches through GATNet('GAT','CIFAR10',5): 128 and 512 graphs at N = 1;
                         data parallel strong (128 graphs split over the ranks) and weak (128 per rank) at N > 1
    "cora"   configs[0]  Cora-shaped graph through GATNet('GAT','Cora',1433), dropout 0.6 active (replicas only: N = 1)
    "heads"  configs[3]  one 50 -> H x 64 layer on the PPI-shaped batch, H = 1, 2, 4, 8, 16 (replicas only: N = 1)
`roofline` is the dominant C-ABI op of a workload's step, timed live with CUDA events around each C-ABI call; `kernels` lists
every op; at N > 1 `check` compares the distributed step-0 loss and gradients with the same global batch run on ONE GPU.
`cpu_baseline` / `--impl reference` time the reference's own GAT.py / GATNet.py (unmodified, staged by
oracle/stage_reference.py under oracle/_ref/, running on oracle/pyg_standin; kind "reference") — or the oracle port when
the staged files are absent (kind "port") — on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
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
ALL_WORKLOADS = ("ppi", "large", "cifar", "cora", "heads")
HEAD_POINTS = (1, 2, 4, 8, 16)          # BASELINE configs[3]
L2_BYTES = 126 << 20


# ------------------------------------------------------------------------------------------------ workloads
def make_workload(name, seed, sample=None, heads=8, num_graphs=128):
    """-> (data, spec, loss_fn, description).  `sample` bounds the CPU legs (a part of the batch / a scaled graph)."""
    import torch.nn.functional as F
    if name == "ppi":
        data = synth.ppi_shaped(seed=seed, keep_graphs=sample)
        spec = synth.PPI_STACK
        loss_fn = lambda out, y: F.binary_cross_entropy_with_logits(out, y)   # noqa: E731
        desc = "PPI-shaped inductive batch (24 graphs, 56944 nodes, 818716 edges, 50 feats, 121 labels), 3-layer GAT 4/4/6 heads x 256/256/121"
    elif name == "large":
        data = synth.powerlaw(seed=seed) if sample is None else synth.powerlaw(seed=seed, num_nodes=2_400_000 // sample,
                                                                               num_edges=62_000_000 // sample)
        spec = synth.LARGE_STACK
        loss_fn = lambda out, y: F.nll_loss(F.log_softmax(out, dim=1), y)     # noqa: E731
        desc = "large power-law graph (2.4M nodes, 62M edges, 100 feats, 47 classes), 3-layer GAT 4 heads x 128"
    elif name == "heads":
        data = synth.ppi_shaped(seed=seed, keep_graphs=sample)
        spec = [(50, 64, heads, True)]
        loss_fn = lambda out, y: out.sum()                                     # noqa: E731
        desc = f"heads sweep point: one layer 50 -> {heads} heads x 64 on the PPI-shaped batch"
    elif name == "cifar":
        # BASELINE configs[2]: run_gnn_benchmark.py:35-66 — GATNet('GAT','CIFAR10',F): conv1 -> elu -> conv2 -> elu ->
        # per-graph mean -> lin1 -> relu -> lin2 -> log_softmax, nll_loss per graph; dropout 0.0 (GATNet.py:19-20)
        data = synth.cifar_shaped(seed=seed, num_graphs=num_graphs if sample is None else sample)
        spec = [(5, 8, 8, True), (64, 8, 8, True)]
        loss_fn = lambda out, y: F.nll_loss(out, y)                            # noqa: E731
        desc = (f"CIFAR10-superpixel-shaped batch ({num_graphs} graphs x ~117 nodes, kNN k=8, 5 feats, 10 classes), "
                "GATNet('GAT','CIFAR10',5)")
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


def working_set_bytes(spec, n, ep, cached):
    return sum(algorithmic_bytes(op, n, ep, f, c, h, concat, li > 0, cached)[0]
               for li, (f, c, h, concat) in enumerate(spec)
               for op in ("b200gat_proj_fwd", "b200gat_edge_fwd", "b200gat_edge_bwd", "b200gat_proj_bwd"))


def describe_config(name, desc, data, spec, world, mode):
    """The `config` object of a line: the WORKLOAD only, so the b200 arm and the reference arm print the same dict."""
    n, ep = int(data.x.shape[0]), layer_edges(data)
    ws = working_set_bytes(spec, n, ep, cached=name != "large")
    flush = ws < 2 * L2_BYTES
    return {"workload": desc, "nodes": n, "input_edges": int(data.edge_index.shape[1]),
            "edges_per_layer_incl_self_loops": ep, "layers": len(spec),
            "step": "zero_grad + fwd + loss + bwd + grad all-reduce (N > 1) + fused Adam (lr 5e-3, wd 5e-4)",
            "parallelism": mode if world > 1 else "single GPU",
            "l2": (f"per-step algorithmic traffic {ws / 1e6:.0f} MB < 2 x L2 (126 MB): L2 flushed (256 MB memset) before every timed step, "
                   "steps timed one by one with CUDA events" if flush else
                   f"inputs larger than L2 (per-step algorithmic traffic {ws / 1e9:.2f} GB vs 126 MB L2); no explicit flush")}, flush


# ------------------------------------------------------------------------------------ algorithmic bytes
def algorithmic_bytes(op, n, ep, f, c, h, concat, need_gx, cached, row_b=4):
    """Bytes one ABI op must move (fp32 values, int32 indices); G = gather multiplicity: N when the gathered operand
    of the largest graph block fits half of L2 ("cached regime": the block-diagonal PPI-shaped batch, blocks <= 15 MB),
    E' otherwise ("streaming regime": the 2.4M-node graph, 4.9 GB of Wh) — SURVEY.md §8d."""
    d = h * c
    d_out = d if concat else c
    g = n if cached else ep
    # row_b = bytes per element of the GATHERED rows: 4 (default) or 2 (bf16 gather mode: the bf16 copies are extra writes)
    if op == "b200gat_proj_fwd":
        return 4 * (n * f + d * f + n * d + 2 * n * h) + (2 * n * d if row_b == 2 else 0), "tensor"
    if op == "b200gat_edge_fwd":
        extra = 4 * n * d if (not concat and h > 1) else 0
        return 4 * ((n + 1) + ep + 2 * n * h + n * d_out) + g * (row_b * d + 4 * h) + extra, "hbm"
    if op == "b200gat_edge_bwd":
        return (4 * ((n + 1) + ep + n * d_out + 3 * n * d + 4 * n * h) + g * (row_b * d_out + 16 * h) +
                (2 * n * d_out if row_b == 2 else 0)), "hbm"
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


def git_sha():
    try:
        return subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip() or None
    except Exception:   # pragma: no cover
        return None


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


# ------------------------------------------------------------------------------------------------ CPU legs
CPU_SAMPLES = {"ppi": 2, "heads": 2, "cifar": 16, "cora": None, "large": 128}   # graphs kept / scale divisor


def cpu_reference_leg(workload, steps, warmup, sample, heads=8):
    """The reference's CPU PyTorch path on all host cores: its own GAT.py / GATNet.py (oracle/_ref or /root/reference, on
    oracle/pyg_standin) when present, else the oracle port (same op sequence, autograd backward).
    -> (edges/s, seconds per step, sample description, kind)."""
    from oracle import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    data, spec, loss_fn, _ = make_workload(workload, 0, sample=sample, heads=heads)
    kind = "reference" if ref_loader.available() else "port"
    torch.manual_seed(0)
    if workload in ("cifar", "cora"):
        ds = "CIFAR10" if workload == "cifar" else "Cora"
        if kind == "reference":
            net = ref_loader.load()[1].GATNet("GAT", ds, data.x.shape[1]).train()
        else:
            from oracle.gat_port import PortGATNet
            net = PortGATNet("GAT", ds, data.x.shape[1]).train()
        model = lambda x, ei: net(data)                                        # noqa: E731
        params = list(net.parameters())
    else:
        if kind == "reference":
            stack = ref_loader.RefStack(spec, dropout=0.0)
        else:
            from oracle.gat_port import PortStack
            stack = PortStack(spec, dropout=0.0)
        model, params = stack, list(stack.parameters())
    opt = torch.optim.Adam(params, lr=5e-3, weight_decay=5e-4)
    ep = layer_edges(data)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(data.x, data.edge_index), data.y)
        loss.backward()
        opt.step()
        return float(loss.detach())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    if workload in ("ppi", "heads"):
        what = "the whole PPI-shaped batch (24 graphs)" if sample is None else f"{data.num_graphs} of 24 graphs of the PPI-shaped batch"
    elif workload == "cifar":
        what = f"{data.num_graphs} graphs of the CIFAR-shaped batch"
    elif workload == "cora":
        what = "the whole Cora-shaped graph"
    else:
        what = "the whole 2.4M-node graph" if sample is None else f"1/{sample}-scale graph from the same power-law generator"
    impl = ("the reference's own GAT.py / GATNet.py (unmodified) on oracle/pyg_standin" if kind == "reference" else
            "oracle/gat_port.py (port of GAT.py: same op sequence, autograd backward)")
    desc = (f"{what}: {data.x.shape[0]} nodes, {ep} edges incl. self loops, {len(spec)} layers, {steps} steps after "
            f"{warmup} warm-up; {impl}")
    return len(spec) * ep / dt, dt, desc, kind


def cpu_baseline_record(workload, heads=8):
    sample = CPU_SAMPLES[workload]
    steps, warm = (2, 1) if workload == "large" else (3, 1)
    val, dt, desc, kind = cpu_reference_leg(workload, steps, warm, sample, heads=heads)
    return {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": desc, "s_per_step": dt}


# ------------------------------------------------------------------------------------------------ output
_REAL_STDOUT = 1


def _emit(line):
    """The ONE JSON line goes to the real stdout; everything else that lands on fd 1 while the bench runs (NCCL prints
    its version banner there) is diverted to stderr so that the line stays machine-readable."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ one GPU workload
class Env:
    def __init__(self, rank, local_rank, world, dev):
        self.rank, self.local_rank, self.world, self.dev = rank, local_rank, world, dev

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = torch.tensor([float(v)], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


class Runner:
    """One workload on this rank's GPU: model, optimizer, host / device buffers and the train step."""

    def __init__(self, name, env, *, heads=8, num_graphs=128, dp="weak", capturable=False, gather_bf16=False):
        from types import SimpleNamespace
        from atmlgraphattentionnetworks_b200.gatnet import GATNet, GATStack
        from atmlgraphattentionnetworks_b200.parallel import GradBucket, shard_graphs
        self.name, self.env, self.heads = name, env, heads
        world, rank, dev = env.world, env.rank, env.dev
        self.partitioned = name == "large" and world > 1
        self.dp = dp if (world > 1 and not self.partitioned) else None
        # graph batches, weak: every rank owns its own batch of independent graphs (seed = rank), model replicated;
        # strong: the seed-0 global batch is split by graph (parallel.shard_graphs);
        # one large graph (N > 1): strong scaling — destination-row partition, Wh all-gather per layer.
        seed = rank if self.dp == "weak" else 0
        data, spec, loss_fn, desc = make_workload(name, seed=seed, heads=heads, num_graphs=num_graphs)
        self.global_data = data
        mode = (f"row{world} (destination-row partition; input features replicated, so layer 1 exchanges nothing forward; all-gather of Wh (layers 2, 3) / gradient rows (all layers), reduce-scatter of g_s_dst, grad all-reduce)"
                if self.partitioned else
                f"dp{world} {self.dp} (independent graphs per rank, one NCCL all-reduce of the persistent gradient arena)")
        self.config, self.flush = describe_config(name, desc, data, spec, world, mode)
        if self.dp == "strong":
            data = synth.select_graphs(data, shard_graphs(data.num_graphs, world, rank))
        self.data, self.spec, self.loss_fn = data, spec, loss_fn
        self.n, self.ep = data.x.shape[0], layer_edges(data)
        torch.manual_seed(0)
        self.is_net = name in ("cifar", "cora")
        if self.is_net:
            self.net = GATNet("GAT", "CIFAR10" if name == "cifar" else "Cora", data.x.shape[1]).to(dev).train()
            batch_d = data.batch.to(dev) if hasattr(data, "batch") else None
            self.model = lambda x, ei: self.net(SimpleNamespace(x=x, edge_index=ei, batch=batch_d, num_graphs=data.num_graphs))
            self.module = self.net
        else:
            self.module = GATStack(spec, dropout=0.0).to(dev)
            self.model = self.module
        if gather_bf16:      # the separately toleranced mode: gathered rows stored as bf16
            from atmlgraphattentionnetworks_b200.gat import set_gather_dtype
            set_gather_dtype(self.module, torch.bfloat16)
        self.gather_bf16 = bool(gather_bf16)
        self.params = list(self.module.parameters())
        # gradients are set to None every step and autograd adopts the kernels' packed output buffers as .grad without a
        # copy; N > 1 (graph batches): those few buffers are all-reduced in ONE NCCL group (all_reduce_packed_grads).
        # The row-partitioned large graph keeps the flat bucket (its stage functions build gradients through torch ops).
        self.bucket = GradBucket(self.params) if self.partitioned else None
        self.arena = None
        if self.dp is not None:          # data parallel: the GAT layers' gradients live in one persistent arena => one all-reduce
            from atmlgraphattentionnetworks_b200.parallel import ArenaExchange
            self.arena = ArenaExchange(self.module)
        self.opt = torch.optim.Adam(self.params, lr=5e-3, weight_decay=5e-4, fused=True,   # run_inductive.py:18-19,65
                                    capturable=capturable)
        self.part = None
        if self.partitioned:
            from atmlgraphattentionnetworks_b200.partition import PartitionedGATStack, build_row_partition
            self.part = build_row_partition(data.edge_index.to(dev), self.n, world, rank)
            self.pmodel = PartitionedGATStack(self.module)
            self.x_h = data.x[self.part.lo:self.part.hi].contiguous().pin_memory()
            self.y_h = data.y[self.part.lo:self.part.hi].contiguous().pin_memory()
            self.ei_h = torch.zeros((2, 0), dtype=torch.int64).pin_memory()    # the partitioned graph is static and resident
            # the static input features are replicated (loaded once, like the graph): layer 1 projects every node on every
            # rank and exchanges nothing in its forward (partition.PartitionedGATFunction, x_full)
            self.x_full_d = data.x.to(dev) if os.environ.get("B200GAT_REPLICATE_X", "1") == "1" else None
            torch.cuda.empty_cache()
        else:
            self.x_h, self.ei_h, self.y_h = data.x.pin_memory(), data.edge_index.pin_memory(), data.y.pin_memory()
        self.x_d, self.ei_d, self.y_d = self.x_h.to(dev), self.ei_h.to(dev), self.y_h.to(dev)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if self.flush else None

    # -------------------------------------------------------------------------------------------- the step
    def forward_backward(self, x, ei, y):
        """zero_grad + forward + loss + backward + gradient exchange; -> loss (this rank's share in partitioned mode)"""
        from atmlgraphattentionnetworks_b200.parallel import all_reduce_packed_grads
        if self.bucket is not None:
            self.bucket.zero()
        else:
            self.opt.zero_grad(set_to_none=True)
        if self.partitioned:
            out = self.pmodel(x, self.part, x_full=self.x_full_d)
            # global mean loss = sum over ranks of (own sum / N); parameter gradients are then SUMMED over ranks
            loss = torch.nn.functional.nll_loss(torch.nn.functional.log_softmax(out, dim=1), y, reduction="sum") / self.n
            loss.backward()
            self.bucket.all_reduce_mean(weight=1.0)
        else:
            loss = self.loss_fn(self.model(x, ei), y)
            loss.backward()
            if self.arena is not None:
                self.arena.all_reduce()
            elif self.env.world > 1:
                all_reduce_packed_grads(self.params)
        return loss

    def train_step(self, x, ei, y):
        loss = self.forward_backward(x, ei, y)
        self.opt.step()
        return loss

    def resident_step(self):
        return self.train_step(self.x_d, self.ei_d, self.y_d)

    def captured_step(self):
        """The train step through the product's whole-step capture (atmlgraphattentionnetworks_b200.capture.CapturedStep):
        static input buffers, the CSR / CSC build of the current edge_index INSIDE the graph, one graph launch per step."""
        from atmlgraphattentionnetworks_b200.capture import CapturedStep
        assert not self.partitioned      # data parallel: the NCCL gradient all-reduce is captured with the step

        def fn(x, edge_index, y):
            return self.train_step(x, edge_index, y).detach()
        # full-graph workloads (run_inductive.py:77: the same data.edge_index every epoch) capture with a static graph —
        # the CSR is built once; the CIFAR-shaped loop (run_gnn_benchmark.py:60-63: a new batch every step) re-ingests its
        # edge_index inside the graph on every replay
        return CapturedStep(fn, dict(x=self.x_d, edge_index=self.ei_d, y=self.y_d), num_nodes=self.n,
                            static_graph=self.name != "cifar")

    def time_captured(self, steps, warmup):
        """-> record: resident replay (static buffers already hold the batch; the graph still re-ingests edge_index) and
        end to end (pinned host -> static buffers, replay, loss -> pinned host, every step)."""
        from atmlgraphattentionnetworks_b200 import _abi
        l0 = _abi.launch_count()
        cap = self.captured_step()
        for _ in range(warmup):
            cap()
        torch.cuda.synchronize()
        launches_total = _abi.launch_count() - l0                 # warm-up runs + ONE capture
        ms = self.timed(cap, steps)
        loss_h = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

        def e2e_run(nsteps):
            pending, losses = None, []
            for k in range(nsteps):
                if cap.static_graph:
                    out = cap(x=self.x_h, y=self.y_h)                         # the graph is resident: features / labels only
                else:
                    out = cap(x=self.x_h, edge_index=self.ei_h, y=self.y_h)   # H2D into the static buffers + replay
                loss_h[k % 2].copy_(out, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[0].synchronize()
                    losses.append(float(pending[1]))
                pending = (ev, loss_h[k % 2])
            pending[0].synchronize()
            losses.append(float(pending[1]))
            assert all(v == v for v in losses)
        e2e_run(3)
        flush, self.flush_buf = self.flush_buf, None
        try:
            ms_e2e = statistics.median(self.timed(lambda: e2e_run(steps), 1) / steps for _ in range(3))
        finally:
            self.flush_buf = flush
        cap.check()
        h2d = self.x_h.numel() * 4 + (0 if cap.static_graph else self.ei_h.numel() * 8) + self.y_h.numel() * self.y_h.element_size()
        per_run = launches_total // (3 + 1)                       # CapturedStep: 3 warm-up runs + the capture
        return {"api": ("atmlgraphattentionnetworks_b200.capture.CapturedStep, static_graph=True (full-graph training: CSR built once, "
                        "train step replayed)" if cap.static_graph else
                        "atmlgraphattentionnetworks_b200.capture.CapturedStep (train step + in-graph CSR/CSC build of the batch's edge_index)"),
                "ms_per_step": ms, "value": self.total_edges() / (ms / 1e3), "unit": UNIT,
                "gpu_launches_per_replay": per_run,
                "e2e": {"ms_per_step": ms_e2e, "value": self.total_edges() / (ms_e2e / 1e3), "unit": UNIT,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "includes": "every step: pinned host -> static device buffers (x, edge_index, y), ONE graph launch (CSR/CSC "
                                    "build + train step), loss D2H into pinned memory read one step behind"}}

    def total_edges(self):
        """edges all ranks process per step (sum over layers)"""
        per_rank = len(self.spec) * self.ep
        if self.partitioned:
            return per_rank
        if self.dp == "strong":
            return len(self.spec) * layer_edges(self.global_data)
        return per_rank * self.env.world

    # -------------------------------------------------------------------------------------------- timing
    def timed(self, fn, steps):
        """ms per step, max over ranks.  Back to back between one event pair — or, for workloads whose working set fits
        L2, step by step with an L2 flush before each (outside the event pair)."""
        env = self.env
        env.barrier()
        if self.flush_buf is None:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                fn()
            e.record()
            env.barrier()
            return env.max_over_ranks(s.elapsed_time(e)) / steps
        pairs = []
        for _ in range(steps):
            self.flush_buf.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            pairs.append((s, e))
        env.barrier()
        return env.max_over_ranks(sum(s.elapsed_time(e) for s, e in pairs)) / steps

    # -------------------------------------------------------------------------------------------- N > 1 correctness
    def check(self):
        """Step-0 loss and gradients of the distributed run against the SAME global batch on one GPU (rank 0 runs every
        rank's share sequentially through the single-GPU path, seed-fixed weights, before any optimizer step)."""
        import torch.distributed as dist
        env = self.env
        world, dev = env.world, env.dev
        loss = self.forward_backward(self.x_d, self.ei_d, self.y_d).detach().clone()
        if self.partitioned:
            dist.all_reduce(loss)                         # sum of the ranks' shares of the global mean
            got = self.bucket.flat.clone()
        else:
            loss = loss / world
            dist.all_reduce(loss)                         # mean of the ranks' (equal-size shard) mean losses
            got = torch.cat([p.grad.flatten() for p in self.params])
        rec = None
        if env.rank == 0:
            from atmlgraphattentionnetworks_b200.graph import GLOBAL_CACHE
            from atmlgraphattentionnetworks_b200.parallel import shard_graphs
            for p in self.params:
                p.grad = None
            want_loss, want = 0.0, None
            if self.partitioned:
                d = self.global_data
                x, ei, y = d.x.to(dev), d.edge_index.to(dev), d.y.to(dev)
                l1 = self.loss_fn(self.module(x, ei), y)
                l1.backward()
                want_loss = float(l1.detach())
                want = torch.cat([p.grad.flatten() for p in self.params]).clone()
                del x, ei, y, l1
            else:
                for r in range(world):
                    if self.dp == "weak":
                        d, _, _, _ = make_workload(self.name, seed=r, heads=self.heads, num_graphs=self.global_data.num_graphs)
                    else:
                        d = synth.select_graphs(self.global_data, shard_graphs(self.global_data.num_graphs, world, r))
                    x, ei, y = d.x.to(dev), d.edge_index.to(dev), d.y.to(dev)
                    if self.is_net:
                        from types import SimpleNamespace
                        out = self.net(SimpleNamespace(x=x, edge_index=ei, batch=d.batch.to(dev), num_graphs=d.num_graphs))
                    else:
                        out = self.module(x, ei)
                    l1 = self.loss_fn(out, y) / world
                    l1.backward()                          # gradients accumulate over the shards
                    want_loss += float(l1.detach())
                want = torch.cat([p.grad.flatten() for p in self.params]).clone()
            for p in self.params:
                p.grad = None
            if self.bucket is not None:                    # re-attach the flat bucket's views
                for p in self.bucket.params:
                    p.grad = self.bucket._slot(p)
            GLOBAL_CACHE.clear()
            torch.cuda.empty_cache()
            loss_rel = abs(float(loss) - want_loss) / max(abs(want_loss), 1e-30)
            grad_rel = float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))
            tol_l, tol_g = (1e-2, 1e-2) if self.gather_bf16 else (1e-5, 2e-5)       # the bf16 gather mode's own tolerance
            rec = {"ok": bool(loss_rel <= tol_l and grad_rel <= tol_g), "loss_distributed": float(loss), "loss_one_gpu": want_loss,
                   "loss_rel_err": loss_rel, "grad_max_rel_err": grad_rel, "tolerance": {"loss_rel": tol_l, "grad_max_rel": tol_g},
                   "what": ("step-0 loss and all parameter gradients after the exchange vs the same global batch on rank 0's GPU alone "
                            "(single-GPU path, same seed-fixed weights)")}
        env.barrier()
        return rec

    # -------------------------------------------------------------------------------------------- e2e
    def time_e2e(self, steps, reps=3):
        """pinned host buffers in, loss out, every step (new tensors => CSR rebuilt every step).  The host->device copy of
        step k+1 is issued on a copy stream while step k computes (what a data loader with a prefetch depth of one does):
        every step still copies its own inputs from pinned host memory inside the timed region, builds its CSR from the
        fresh edge_index, and reads its loss back."""
        from atmlgraphattentionnetworks_b200.graph import GLOBAL_CACHE, build_csr
        copy_stream = torch.cuda.Stream()
        # graph batches are ingested WITHOUT a host synchronisation (graph.build_csr(sync=False): no status read-back, the
        # index check is deferred to one lazy check per run) so the host never blocks on the copy stream; the power-law
        # graph needs its degree classes (hub / giant rows), i.e. the synchronous build
        lazy = self.name != "large"
        built = []
        # two static sets of device buffers (no per-step allocation on the copy stream: cross-stream frees make the caching
        # allocator synchronise); the in-place copy bumps edge_index's version counter, so the graph cache misses and the
        # CSR is rebuilt for every batch
        slots = [(torch.empty_like(self.x_d), torch.empty_like(self.ei_d), torch.empty_like(self.y_d)) for _ in range(2)]
        consumed = [None, None]          # event: the step that read slot k has finished
        loss_h = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

        def upload(k):
            """batch k: pinned host -> device slot k % 2, then graph ingestion (GAT.py:38 + the CSR / CSC / degree-class
            build: b200gat_csr_build, b200gat_hub_rows, one status read) — on the copy stream, while step k - 1 computes."""
            x, ei, y = slots[k % 2]
            with torch.cuda.stream(copy_stream):
                if consumed[k % 2] is not None:
                    copy_stream.wait_event(consumed[k % 2])
                x.copy_(self.x_h, non_blocking=True)
                ei.copy_(self.ei_h, non_blocking=True)
                y.copy_(self.y_h, non_blocking=True)
                if not self.partitioned:
                    if lazy:
                        built.append(GLOBAL_CACHE.put(ei, self.n, build_csr(ei, self.n, sync=False)))
                        del built[:-2]
                    else:
                        GLOBAL_CACHE.get(ei, self.n)   # the layers of step k find this batch's CSR in the cache
                done = torch.cuda.Event()
                done.record(copy_stream)
            return (x, ei, y), done

        def e2e_run(nsteps):
            nxt = upload(0)
            pending = None                       # (event, pinned scalar) of the previous step's loss
            losses = []
            for k in range(nsteps):
                (x, ei, y), done = nxt
                torch.cuda.current_stream().wait_event(done)
                loss = self.train_step(x, ei, y)
                loss_h[k % 2].copy_(loss.detach(), non_blocking=True)      # device -> pinned host, read one step later
                ev = torch.cuda.Event()
                ev.record()
                consumed[k % 2] = ev
                if k + 1 < nsteps:
                    nxt = upload(k + 1)          # overlaps step k on the GPU (the host blocks on the ingestion status read)
                if pending is not None:
                    pending[0].synchronize()
                    losses.append(float(pending[1]))
                pending = (ev, loss_h[k % 2])
            pending[0].synchronize()
            losses.append(float(pending[1]))
            assert len(losses) == nsteps and all(v == v for v in losses)
            if built:
                built[-1].check()            # the deferred index check of the last ingested batch (one D2H word per run)
        e2e_run(3)                           # warm-up: also fills the copy stream's allocator pool (CSR arrays, sort workspace)
        flush, self.flush_buf = self.flush_buf, None     # e2e streams fresh inputs from the host every step: no flush needed
        try:
            ms = statistics.median(self.timed(lambda: e2e_run(steps), 1) / steps for _ in range(reps))
        finally:
            self.flush_buf = flush
        h2d = self.x_h.numel() * 4 + self.ei_h.numel() * 8 + self.y_h.numel() * self.y_h.element_size()
        del slots
        return ms, h2d

    # -------------------------------------------------------------------------------------------- per-op breakdown
    def per_op(self, reps, traffic_name):
        from atmlgraphattentionnetworks_b200 import _abi
        from atmlgraphattentionnetworks_b200.graph import GLOBAL_CACHE
        GLOBAL_CACHE.clear()
        self.resident_step()
        _abi.timing = []
        for _ in range(reps):
            self.resident_step()
        torch.cuda.synchronize()
        per_raw = {}
        for name, geom, s, e in _abi.timing:
            per_raw.setdefault((name, geom), []).append(s.elapsed_time(e))
        _abi.timing = None
        per, coll = {}, {}
        for (name, geom), ts in per_raw.items():
            if not name.startswith("b200gat_"):          # collectives / barriers of the partitioned path
                per_step = sum(ts) / reps
                coll[name] = coll.get(name, 0.0) + per_step
                continue
            # staged backward (partitioned mode): prep + csc + finish = edge_bwd
            key = ("b200gat_edge_bwd" if name.startswith("b200gat_edge_bwd") else name, geom)
            per[key] = [a + b for a, b in zip(per[key], ts)] if key in per else list(ts)
        n_loc = self.part.n_own if self.partitioned else self.n
        ep_loc = int(self.part.col.numel() + self.part.crow.numel()) // 2 if self.partitioned else self.ep
        peaks = measured_peaks()
        cached = self.name != "large"
        kernels = []
        for li, (f, c, h, concat) in enumerate(self.spec):
            for op in ("b200gat_proj_fwd", "b200gat_edge_fwd", "b200gat_edge_bwd", "b200gat_proj_bwd"):
                ts = per.get((op, (f, c, h, bool(concat))), [])
                if not ts:
                    continue
                if len(ts) > reps:      # two layers of one geometry (CIFAR conv... never equal today) — split evenly
                    ts = ts[:reps]
                ms = statistics.median(ts)
                need_gx = li > 0
                nbytes, bound = algorithmic_bytes(op, n_loc, ep_loc, f, c, h, concat, need_gx, cached=cached,
                                                  row_b=2 if self.gather_bf16 else 4)
                rec = {"op": op, "layer": li, "geom": f"{f}->{h}x{c}{'cat' if concat else 'mean'}", "ms": ms,
                       "alg_bytes": nbytes, "GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peaks["hbm"]}
                fl = gemm_flops(op, n_loc, f, c, h, need_gx)
                if fl:
                    rec["TFLOPs"] = fl / ms / 1e9
                kernels.append(rec)
        dom = max(kernels, key=lambda r: r["ms"]) if kernels else None
        # measured DRAM traffic of each op: one committed `ncu --set full` capture of this workload (tools/ncu_traffic.py),
        # stamped with the git SHA of the kernels it was taken from
        traffic, traffic_src = {}, None
        tpath = os.path.join(ROOT, "profiles", f"traffic_{traffic_name}.json")
        if os.path.isfile(tpath) and not self.partitioned:
            tj = json.load(open(tpath))
            traffic = tj.get("ops", {})
            traffic_src = {"file": f"profiles/traffic_{traffic_name}.json", "capture": tj.get("source"), "git_sha": tj.get("git_sha")}
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
                            "traffic": dom["dram_traffic_bytes"], "alg_bytes": dom["alg_bytes"],
                            "note": "3 fp16 MMA passes per fp32-accurate product: frac is bounded by 0.33 (DESIGN.md §4.3)"}
            roofline.update({"kernel": f"{dom['op']} layer {dom['layer']} ({dom['geom']})", "ms": dom["ms"],
                             "peak_source": peaks["source"] + " (of measured)", "traffic_source": traffic_src})
            if cached and roofline["bound"] == "hbm":
                # SURVEY.md §8d caveat 2: in the cached regime the HBM count is the compulsory traffic; every edge still
                # pulls its row through L2 -> SM (4.4x the HBM bytes on the PPI-shaped layers), which bounds these kernels
                roofline["regime"] = "cached: gathered rows are L2-resident, the kernel is bound by L2->SM gather traffic (DESIGN.md §5)"
            elif roofline["bound"] == "hbm":
                roofline["regime"] = "streaming: every gathered row comes from HBM"
        # SURVEY.md §8d caveat 2: whatever the HBM count says, every edge pulls its gathered row through L2 -> SM:
        # 4 (D + H) bytes in the forward, 4 (D_out + 4 H) in the backward.  The "gather service rate" is that traffic over the
        # edge-kernel time; the bare 128-bit gather loop of tools/microbench/gather_bench.cu reaches 18 TB/s on the L2-resident
        # PPI-shaped batch and 6.0-6.7 TB/s from HBM on the power-law graph — the bound of the cached regime
        for rec in kernels:
            li = rec["layer"]
            f, c, h, concat = self.spec[li]
            d = h * c
            d_out = d if concat else c
            rb = 2 if self.gather_bf16 else 4
            if rec["op"] == "b200gat_edge_fwd":
                rec["gather_bytes"] = ep_loc * (rb * d + 4 * h)
            elif rec["op"] == "b200gat_edge_bwd":
                rec["gather_bytes"] = ep_loc * (rb * d_out + 16 * h)
            if "gather_bytes" in rec:
                rec["gather_TBps"] = rec["gather_bytes"] / rec["ms"] / 1e9
        edge_ms = sum(r["ms"] for r in kernels if r["op"].startswith("b200gat_edge"))
        edge_bytes = sum(r["alg_bytes"] for r in kernels if r["op"].startswith("b200gat_edge"))
        edge_phase = {"per_rank": self.partitioned, "ms": edge_ms, "alg_bytes": edge_bytes,
                      "GBps": edge_bytes / edge_ms / 1e6 if edge_ms else None,
                      "frac_of_measured_hbm": edge_bytes / edge_ms / 1e6 / peaks["hbm"] if edge_ms else None,
                      "frac_of_nominal_8TBps": edge_bytes / edge_ms / 1e6 / 8000.0 if edge_ms else None,
                      "edges_per_s": len(self.spec) * ep_loc / (edge_ms / 1e3) if edge_ms else None}
        gbytes = sum(r.get("gather_bytes", 0) for r in kernels)
        if edge_ms and gbytes:
            ref_rate = 18.0 if cached else 6.5      # TB/s: the bare gather loop on this regime (gather_bench.cu, round 1)
            edge_phase.update({"gather_bytes_through_l2": gbytes, "gather_service_TBps": gbytes / edge_ms / 1e9,
                               "bare_gather_loop_TBps": ref_rate,
                               "frac_of_bare_gather_loop": gbytes / edge_ms / 1e9 / ref_rate,
                               "gather_note": ("cached regime: rows come from L2; the whole op (prep / finish streaming passes "
                                               "included) is compared with a loop that does nothing but the gathers"
                                               if cached else "streaming regime: rows come from HBM")})
        return kernels, roofline, edge_phase, (coll or None)

    def close(self):
        from atmlgraphattentionnetworks_b200.graph import GLOBAL_CACHE
        GLOBAL_CACHE.clear()
        for k in list(self.__dict__):
            if k not in ("name", "env"):
                delattr(self, k)
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def run_workload(name, env, steps, warmup, args, *, heads=8, num_graphs=128, dp="weak", e2e=True, cpu=True, breakdown=True,
                 cuda_graph=False, traffic_name=None, captured=False, gather_bf16=False):
    """-> the record of one workload (all ranks take part; only rank 0's record is complete)."""
    from atmlgraphattentionnetworks_b200 import _abi
    t_wall = time.perf_counter()
    run = Runner(name, env, heads=heads, num_graphs=num_graphs, dp=dp, capturable=cuda_graph or captured, gather_bf16=gather_bf16)
    rec = {"config": run.config, "n_gpus": env.world, "steps": steps, "warmup": warmup,
           "gather_dtype": "bf16 (gathered rows stored as bf16, fp32 arithmetic; tolerance 1e-2, tests/test_gpu_parity.py)" if gather_bf16 else "f32",
           "scaling": "strong" if (run.partitioned or run.dp == "strong") else "weak"}
    if env.world > 1:
        rec["check"] = run.check()
    for _ in range(warmup):
        run.resident_step()
    resident = run.resident_step
    launches_per_replay = None
    if cuda_graph:
        # whole-step capture through the product API: zero_grad + forward + loss + backward + fused
        # Adam — and the CSR / CSC build of the batch's edge_index — become ONE graph launch
        assert env.world == 1, "--cuda-graph is a single-GPU mode"
        l0 = _abi.launch_count()
        cap = run.captured_step()
        launches_per_replay = (_abi.launch_count() - l0) // 4      # 3 warm-up runs + the capture
        resident = cap
        for _ in range(3):
            resident()
    launches0 = _abi.launch_count()
    with ClockSampler(env.local_rank) as clocks:
        ms_step = run.timed(resident, steps)
    launches = _abi.launch_count() - launches0
    if launches_per_replay is not None:
        launches = launches_per_replay * steps
    rec.update({"ms_per_step": ms_step, "value": run.total_edges() / (ms_step / 1e3), "unit": UNIT, "gpu_launches": launches,
                "cuda_graph": bool(cuda_graph), "clocks": clocks.summary()})
    if args.profile:
        run.close()
        return rec
    if e2e:
        e2e_steps = max(steps // 2, 3)
        ms_e2e, h2d = run.time_e2e(e2e_steps)
        rec["e2e"] = {"value": run.total_edges() / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                      "includes": ("every step: H2D of its batch from pinned host and graph ingestion (CSR/CSC build, 2 radix sorts; "
                                   "sync-free for graph batches, with degree classes + status read for the power-law graph) on a copy "
                                   "stream one batch ahead, train step, loss D2H into pinned memory read one step behind"
                                   if not run.partitioned else
                                   "every step: H2D of the own rows of x / y from pinned host, train step, loss D2H (the partitioned "
                                   "graph is static and resident)")}
    if captured and not run.partitioned:
        try:
            rec["captured"] = run.time_captured(steps, warmup)
        except Exception as exc:      # e.g. an NCCL build that cannot be captured: the eager numbers stand
            if env.world == 1:
                raise
            rec["captured"] = {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}
            torch.cuda.synchronize()
    if breakdown:
        kernels, roofline, edge_phase, coll = run.per_op(5, traffic_name or name)
        rec.update({"roofline": roofline, "edge_phase": edge_phase, "kernels": kernels})
        if coll is not None:
            rec["collectives_ms_per_step"] = coll
            rec["collectives_note"] = ("CUDA events around each collective on the compute stream (serialised with the kernels: "
                                       "the time is fully exposed)")
        ksum = sum(r["ms"] for r in kernels)
        rec["abi_ops_ms_sum"] = ksum
        rec["step_over_abi_ops"] = ms_step / ksum if ksum else None
        if "captured" in rec:
            rec["captured"]["step_over_abi_ops"] = rec["captured"]["ms_per_step"] / ksum if ksum else None
    spec_n = (len(run.spec), run.ep)
    run.close()
    if cpu and env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        rec["cpu_baseline"] = cpu_baseline_record(name, heads=heads)
    rec["wall_s"] = time.perf_counter() - t_wall
    log(f"{name}{'' if name != 'heads' else heads}: {ms_step:.3f} ms/step, {rec['value'] / 1e6:.1f} M edges/s "
        f"({spec_n[0]} layers x {spec_n[1]} edges), wall {rec['wall_s']:.1f} s")
    return rec


# ------------------------------------------------------------------------------------------------ main
def reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path on the b200 arm's config (rank 0 only)."""
    name = "ppi" if args.workload == "all" else args.workload
    sample = None if name in ("ppi", "cifar", "cora", "heads") else CPU_SAMPLES[name]   # the large graph cannot run at full size
    if args.cpu_sample_graphs and name in ("ppi", "heads"):
        sample = args.cpu_sample_graphs
    data, spec, _, desc = make_workload(name, 0, heads=args.heads)
    world = max(args.gpus, 1)
    mode = f"dp{world} weak (independent graphs per rank, one NCCL all-reduce of the persistent gradient arena)"
    config, _ = describe_config(name, desc, data, spec, world, mode)
    del data
    val, dt, sdesc, kind = cpu_reference_leg(name, args.steps, args.warmup, sample, heads=args.heads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sdesc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


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
    ap.add_argument("--workload", default="all", choices=["all"] + list(ALL_WORKLOADS))
    ap.add_argument("--heads", type=int, default=8, help="--workload heads: number of heads (x 64 channels); 0 = the whole sweep")
    ap.add_argument("--graphs", type=int, default=128, help="--workload cifar: graphs per batch")
    ap.add_argument("--dp", default="weak", choices=["weak", "strong"], help="--workload cifar at N > 1")
    ap.add_argument("--gather-dtype", default="f32", choices=["f32", "bf16"],
                    help="storage of the rows the edge kernels gather: f32 (default, 1e-5 parity) or bf16 (tolerance 1e-2)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="capture the resident train step in a CUDA graph and time replays (launch-bound workloads)")
    ap.add_argument("--cpu-sample-graphs", type=int, default=0,
                    help="--impl reference on ppi / heads: keep only this many of the 24 graphs (0 = the whole batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="resident leg only (for ncu runs): warm-up + steps, minimal JSON")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the GAT hot path has no CPU fallback); use --impl reference for the CPU leg")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus} (launch with torchrun)"
    env = Env(rank, local_rank, world, dev)
    warmup = max(args.warmup, 3)
    everything = args.workload == "all"
    head_name = "ppi" if everything else args.workload

    if args.profile:
        rec = run_workload(head_name, env, args.steps, warmup, args, heads=args.heads or 8, num_graphs=args.graphs, dp=args.dp,
                           cuda_graph=args.cuda_graph)
        if rank == 0:
            _emit({"profile_run": True, "ms_per_step": rec["ms_per_step"], "gpu_launches": rec["gpu_launches"]})
        if world > 1:
            dist.destroy_process_group()
        return 0

    def heads_sweep():
        points = []
        for h in HEAD_POINTS:
            r = run_workload("heads", env, args.steps, warmup, args, heads=h, e2e=False, traffic_name=f"heads{h}", captured=True)
            cp = r.get("captured") or {}
            pt = {"heads": h, "ms_per_step": r["ms_per_step"], "value": r["value"], "abi_ops_ms_sum": r.get("abi_ops_ms_sum"),
                  "step_over_abi_ops": r.get("step_over_abi_ops"), "gpu_launches": r["gpu_launches"],
                  "captured_ms_per_step": cp.get("ms_per_step"), "captured_value": cp.get("value"),
                  "captured_step_over_abi_ops": cp.get("step_over_abi_ops"),
                  "edge_phase_frac_of_measured_hbm": (r.get("edge_phase") or {}).get("frac_of_measured_hbm"),
                  "kernels": [{k: v for k, v in kr.items() if k in ("op", "ms", "GBps", "frac_hbm", "TFLOPs")} for kr in r.get("kernels", [])],
                  "cpu_baseline": r.get("cpu_baseline")}
            points.append(pt)
        return {"config": {"workload": "heads sweep as in run_heads_experiment.py: one layer 50 -> H x 64 on the PPI-shaped batch, H = 1, 2, 4, 8, 16"},
                "unit": UNIT, "points": points}

    if head_name == "heads" and not args.heads:
        head = heads_sweep()
        head.update({"ms_per_step": head["points"][-1]["ms_per_step"], "value": head["points"][-1]["value"], "gpu_launches":
                     head["points"][-1]["gpu_launches"], "n_gpus": world, "steps": args.steps, "warmup": warmup, "scaling": "weak"})
    else:
        head = run_workload(head_name, env, args.steps, warmup, args, heads=args.heads or 8, num_graphs=args.graphs, dp=args.dp,
                            cuda_graph=args.cuda_graph, captured=(head_name != "large"),
                            gather_bf16=args.gather_dtype == "bf16")
    subs = {}
    if everything:
        sub_steps = min(args.steps, 10)
        subs["large"] = run_workload("large", env, sub_steps, 3, args)
        # the separately toleranced bf16-gather mode on the two roofline workloads (NOT the parity path; see gather_dtype)
        subs["bf16_gather"] = {
            "note": "gathered rows (Wh forward, gradient rows backward, and the all-gathered copies at N > 1) stored as bf16; "
                    "fp32 arithmetic; tolerance 1e-2 instead of 1e-5 (tests/test_gpu_parity.py::test_bf16_gathered_rows_*)",
            "large": run_workload("large", env, sub_steps, 3, args, gather_bf16=True, e2e=False, cpu=False, traffic_name="large_bf16")}
        # (not run on the PPI-shaped batch: with L2-resident rows the edge kernels are issue-bound, and the 8-byte loads +
        #  conversions of the bf16 rows make them SLOWER — measured edge_fwd 0.257 -> 0.297 ms)
        if world == 1:
            subs["cifar"] = run_workload("cifar", env, args.steps, warmup, args, num_graphs=128, captured=True)
            subs["cifar"]["batch512"] = run_workload("cifar", env, args.steps, warmup, args, num_graphs=512, cpu=False,
                                                     traffic_name="cifar512", captured=True)
            subs["cora"] = run_workload("cora", env, args.steps, warmup, args, captured=True)
            subs["heads"] = heads_sweep()
        else:
            subs["cifar"] = {"dp_strong": run_workload("cifar", env, args.steps, warmup, args, num_graphs=128, dp="strong", captured=True),
                             "dp_weak": run_workload("cifar", env, args.steps, warmup, args, num_graphs=128, dp="weak", captured=True)}
            subs["cora"] = subs["heads"] = {"skipped": "replicas only: the workload does not shard (DESIGN.md §6); see the N = 1 line"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": head.get("scaling", "weak"),
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "git_sha": git_sha()}
    for k in ("config", "e2e", "gpu_launches", "roofline", "edge_phase", "kernels", "abi_ops_ms_sum", "step_over_abi_ops",
              "collectives_ms_per_step", "collectives_note", "check", "cpu_baseline", "clocks", "cuda_graph", "points", "captured"):
        if k in head:
            line[k] = head[k]
    line.update(subs)
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
