"""Kernel / copy timeline of one bench step (torch profiler, CUDA activities): start, gap to the previous activity,
duration, name.  Gaps under the profiler are longer than in a plain run; the ORDER and the durations are what to read.

    python profiles/timeline.py c2 1000000 [device|e2e]
"""
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
mode = sys.argv[3] if len(sys.argv) > 3 else "device"
wl = bench.WORKLOADS[sys.argv[1]](int(sys.argv[2]), 0, dev)
step = wl.step_device if mode == "device" else wl.step_e2e
for _ in range(5):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    ts.append(1e3 * (time.perf_counter() - t0))
print("wall ms per step:", " ".join("%.2f" % t for t in ts))
with profile(activities=[ProfilerActivity.CUDA]) as warm:  # profiler start-up costs land here
    step()
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = t0
busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    busy += d
    print("%8.0f us  gap %6.0f  dur %7.0f  %s" % (e.time_range.start - t0, gap, d, e.name[:70]))
    prev_end = max(prev_end, e.time_range.end)
print("span %.0f us, sum of durations %.0f us" % (prev_end - t0, busy))
