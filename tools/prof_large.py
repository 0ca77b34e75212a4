"""torch.profiler kernel table of one resident train step of a bench workload (finds the non-library time).
usage: python tools/prof_large.py <workload>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "large"
env = bench.Env(0, 0, 1, torch.device("cuda", 0))
torch.cuda.set_device(0)
run = bench.Runner(name, env)
for _ in range(3):
    run.resident_step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        run.resident_step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 2e3:.2f} ms per step")
for e in rows[:28]:
    print(f"{e.device_time_total / 2e3:9.3f} ms  x{e.count // 2:<4d} {e.key[:110]}")
