"""Kernel timeline of one CUDA-graph replay of the bs32 train step (CUPTI through torch.profiler; no nsys in the
image).  Prints GPU busy time (union over streams), the concurrency profile, per-kernel totals and the largest idle
gaps, and writes the raw kernel list to gpurun_out/timeline.csv."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from shg_vqa_b200.trainer import HotPathTrainer

dev = torch.device("cuda:0")
drop = "--no-dropout" not in sys.argv
tr = HotPathTrainer(dev, t_total=100000, dropout=drop)
tr.init_weights()
h_inp, h_batch = bench.make_batch(1234, pin=False)
inp = {k: v.to(dev) for k, v in h_inp.items()}
batch = {k: v.to(dev) for k, v in h_batch.items()}
tr.capture(inp, batch)
for _ in range(3):
    tr.train_step_graph()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_step_graph()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
t0 = ks[0][0]
t1 = max(k[1] for k in ks)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/timeline.csv", "w") as f:
    f.write("start_us,dur_us,name\n")
    for s, e, n in ks:
        f.write("%.2f,%.2f,%s\n" % (s - t0, e - s, n.split("(")[0][:60]))
print("kernels %d  span %.1f us  sum of durations %.1f us" % (len(ks), t1 - t0, sum(e - s for s, e, _ in ks)))
# sweep line: concurrency profile
pts = sorted([(s, 1) for s, e, _ in ks] + [(e, -1) for s, e, _ in ks])
conc, last, hist = 0, t0, {}
gaps = []
for t, d in pts:
    if t > last:
        hist[conc] = hist.get(conc, 0.0) + (t - last)
        if conc == 0:
            gaps.append((t - last, last - t0))
    conc += d
    last = t
print("time at concurrency level (us):", {k: round(v, 1) for k, v in sorted(hist.items())})
print("idle gaps: total %.1f us; largest:" % hist.get(0, 0.0), [(round(g, 1), round(at, 1)) for g, at in sorted(gaps, reverse=True)[:12]])
agg = {}
for s, e, n in ks:
    n = n.split("(")[0][:48]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += e - s
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-50s x%4d %9.1f us  avg %7.1f" % (n, c, t, t / c))
