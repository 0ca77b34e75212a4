"""Debug harness: call the projection ABI directly and compare with torch fp64 matmuls."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from atmlgraphattentionnetworks_b200 import _abi
from atmlgraphattentionnetworks_b200.gat import _layer_struct, _workspace

def run(n, f, c, h, mode="rand"):
    lib = _abi.lib(); dev = torch.device("cuda:0")
    layer = _layer_struct(f, c, h, True); cp = layer.c_pad; dp = h * cp
    torch.manual_seed(0)
    if mode == "ones":
        gt = torch.ones(n, dp, device=dev); x = torch.ones(n, f, device=dev)
    else:
        gt = torch.randn(n, dp, device=dev); x = torch.randn(n, f, device=dev)
    w = torch.randn(dp, f, device=dev)
    gx = torch.empty(n, f, device=dev); gw = torch.full((dp, f), -7.0, device=dev)
    wsb = int(lib.b200gat_proj_bwd_workspace_bytes(ctypes.byref(layer), n)); ws = _workspace(wsb, dev)
    pb = _abi.ProjBwdArgs(layer, n, gt.data_ptr(), x.data_ptr(), f, w.data_ptr(), gx.data_ptr(), f, gw.data_ptr(), ws.data_ptr(), wsb)
    rc = lib.b200gat_proj_bwd(ctypes.byref(pb), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("rc", rc, _abi.last_error() if rc else "", "ws", wsb)
    ref_gw = (gt.double().t() @ x.double()); ref_gx = gt.double() @ w.double()
    e1 = float((gw.double() - ref_gw).abs().max() / ref_gw.abs().max()); e2 = float((gx.double() - ref_gx).abs().max() / ref_gx.abs().max())
    print(f"n={n} f={f} c={c} h={h} {mode}: gW err {e1:.3e}  gX err {e2:.3e}")
    if e1 > 1e-4:
        print("gw[0,:8]", gw[0, :8].tolist()); print("ref   ", ref_gw[0, :8].tolist())
        print("gw[:8,0]", gw[:8, 0].tolist()); print("ref   ", ref_gw[:8, 0].tolist())
        print("nonzero frac", float((gw != 0).float().mean()), "gw abs max", float(gw.abs().max()))

for args in [(512, 64, 32, 4, "ones"), (512, 64, 32, 4, "rand"), (3000, 50, 256, 4, "rand"), (2048, 256, 64, 4, "rand"), (1000, 1024, 256, 4, "rand")]:
    run(*args)
