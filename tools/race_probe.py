"""Is the eager multi-stream backward deterministic?  One GPU.  For each batch size: gradient with every launch on
one stream (use_streams=False) as the reference, then repeated multi-stream runs with the allocator's free blocks
poisoned by NaN in between (an uninitialised read shows as NaN, a stream race as run-to-run noise).
python tools/race_probe.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shg_vqa_b200.trainer import HotPathTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
sizes = [int(x) for x in sys.argv[1:]] or [4, 16, 32]


def poison():
    junk = [torch.full((n,), float("nan"), device=dev) for n in (1 << 28, 1 << 26, 1 << 24, 1 << 22, 1 << 20, 1 << 18)]
    small = [torch.full((1 << 12,), float("nan"), device=dev) for _ in range(256)]
    torch.cuda.synchronize()
    del junk, small


def grads(tr, inp, bat):
    tr.store.grad.zero_()
    out, ctx = tr.engine.forward(inp, train=True)
    _, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, bat)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    return tr.store.grad[tr.store.live_extent].clone()


def per_group(st, a, b):
    lb = st.live_begin
    res = []
    for name, sls in st.dp_groups.items():
        num = den = 0.0
        for sl in sls:
            x, y = a[sl.start - lb:sl.stop - lb].double(), b[sl.start - lb:sl.stop - lb].double()
            num += float((x - y).pow(2).sum()); den += float(y.pow(2).sum())
        res.append("%s %.1e" % (name, (num / max(den, 1e-300)) ** 0.5))
    return "  ".join(res)


for B in sizes:
    h_inp, h_bat = bench.make_batch(4321, pin=False, batch=B)
    inp = {k: v.to(dev) for k, v in h_inp.items()}
    bat = {k: v.to(dev) for k, v in h_bat.items()}
    tr = HotPathTrainer(dev, world_size=1, dropout=False)
    tr.init_weights(seed=77)
    tr.engine.use_streams = False
    ref = grads(tr, inp, bat)
    ref2 = grads(tr, inp, bat)
    print("B=%d one stream, run 2 vs run 1: rel-L2 %.3e" % (B, float((ref2.double() - ref.double()).norm() / ref.double().norm())), flush=True)
    tr.engine.use_streams = True
    for it in range(4):
        poison()
        g = grads(tr, inp, bat)
        bad = int((~torch.isfinite(g)).sum())
        rel = float((g.double() - ref.double()).norm() / ref.double().norm())
        print("B=%d streams run %d: non-finite %d  rel-L2 vs one-stream %.3e   %s" % (B, it, bad, rel, per_group(tr.store, g, ref)), flush=True)
    del tr
    torch.cuda.empty_cache()
