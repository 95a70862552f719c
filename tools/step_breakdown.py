"""Per-shape kernel time of one bs32 train step (CUDA events per launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shg_vqa_b200 import _lib
from shg_vqa_b200.trainer import HotPathTrainer
dev = torch.device("cuda:0")
tr = HotPathTrainer(dev, t_total=100000)
tr.init_weights()
h_inp, h_batch = bench.make_batch(1234, pin=False)
inp = {k: v.to(dev) for k, v in h_inp.items()}
batch = {k: v.to(dev) for k, v in h_batch.items()}
for _ in range(3):
    tr.train_step(inp, batch)
torch.cuda.synchronize()
prof = _lib.Profiler(time_events=True); _lib.profiler = prof
tr.train_step(inp, batch)
agg = prof.by_tag()
_lib.profiler = None
tot = sum(a["ms"] for a in agg.values())
print("total %.2f ms" % tot)
for (name, tag), a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:int(sys.argv[1]) if len(sys.argv) > 1 else 45]:
    rate = a["work"] / (a["ms"] * 1e-3) / 1e12
    print("%-18s %-28s x%3d  %8.1f us  %5.1f%%  avg %7.1f us  %7.1f T/s" % (name, tag, a["calls"], a["ms"] * 1e3, 100 * a["ms"] / tot, a["ms"] * 1e3 / a["calls"], rate))
