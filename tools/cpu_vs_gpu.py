"""Is the step CPU(launch)-bound?  CPU enqueue time vs GPU time per step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shg_vqa_b200.trainer import HotPathTrainer
dev = torch.device("cuda:0")
tr = HotPathTrainer(dev, t_total=100000)
tr.init_weights()
if len(sys.argv) > 1 and sys.argv[1] == "nostreams":
    tr.engine.use_streams = False
h_inp, h_batch = bench.make_batch(1234, pin=False)
inp = {k: v.to(dev) for k, v in h_inp.items()}
batch = {k: v.to(dev) for k, v in h_batch.items()}
for _ in range(5):
    tr.train_step(inp, batch)
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(n):
    tr.train_step(inp, batch)
t_cpu = time.perf_counter() - t0
e.record()
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("streams=%s  cpu enqueue %.2f ms/step   gpu (events) %.2f ms/step   wall %.2f ms/step" % (
    tr.engine.use_streams, 1e3 * t_cpu / n, s.elapsed_time(e) / n, 1e3 * t_all / n))

# CUDA-graph replay
tr.capture(inp, batch)
for _ in range(3):
    tr.train_step_graph()
torch.cuda.synchronize()
t0 = time.perf_counter()
s.record()
for _ in range(n):
    tr.train_step_graph()
t_cpu = time.perf_counter() - t0
e.record()
torch.cuda.synchronize()
print("graph replay: cpu enqueue %.2f ms/step   gpu (events) %.2f ms/step   loss %s" % (1e3 * t_cpu / n, s.elapsed_time(e) / n, tr.losses_from_scalars(tr.static_scal)))
