"""Per-shape kernel time of the given entry points inside one eager bs32 training step (CUPTI durations joined to the
calls through the launch order).  Note: eager launches leave gaps, so L2 residency differs a little from the replayed
graph; shares are what matter.   python tools/shape_times.py shg_layernorm_bwd_grouped shg_colsum_accum ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shg_vqa_b200.trainer import HotPathTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
tr = HotPathTrainer(dev, t_total=100000, warmup=0.1, lr=1e-5, dropout=True, dropout_seed=20261018)
tr.init_weights(seed=9595)
h_inp, h_batch = bench.make_batch(1234, pin=False)
inp = {k: v.to(dev) for k, v in h_inp.items()}
batch = {k: v.to(dev) for k, v in h_batch.items()}
for _ in range(3):
    tr.train_step(inp, batch)
torch.cuda.synchronize()
for entry in sys.argv[1:] or ["shg_layernorm_bwd_grouped"]:
    agg, err = bench.shape_table_from_eager(tr, inp, batch, entry)
    if err:
        print(entry, err)
        continue
    tot = sum(a["us"] for a in agg.values())
    print("%s: %.3f ms in %d calls" % (entry, tot * 1e-3, sum(a["calls"] for a in agg.values())))
    for tag, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        unit = 1e6 if a["kind"] == "tensor" else 1e3
        print("   %-28s calls %3d  avg %7.1f us  total %7.1f us  %s" % (
            tag, a["calls"], a["us"] / a["calls"], a["us"],
            ("%.0f %s" % (a["work"] / a["us"] / unit, "TFLOP/s" if a["kind"] == "tensor" else "GB/s")) if a["work"] else ""))
