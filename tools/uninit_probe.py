"""Does any kernel of the step read memory nobody wrote?  One GPU.  torch.empty / torch.empty_like are replaced by
fills (0, a large finite value, NaN) for the duration of one step; a step whose results depend on the fill reads
uninitialised memory.  Prints which saved tensors / parameter gradients change.
python tools/uninit_probe.py [B] [streams]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shg_vqa_b200.trainer import HotPathTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
streams = len(sys.argv) > 2 and sys.argv[2] == "streams"
real_empty, real_empty_like = torch.empty, torch.empty_like
fill = [None]


def empty(*a, **kw):
    t = real_empty(*a, **kw)
    if fill[0] is not None and t.is_cuda and t.is_floating_point():
        t.fill_(fill[0])
    return t


def empty_like(x, **kw):
    t = real_empty_like(x, **kw)
    if fill[0] is not None and t.is_cuda and t.is_floating_point():
        t.fill_(fill[0])
    return t


torch.empty, torch.empty_like = empty, empty_like


def flat(x, path, acc):
    if torch.is_tensor(x):
        acc.append((path, x))
    elif isinstance(x, dict):
        for k, v in x.items():
            flat(v, path + "." + str(k), acc)
    elif isinstance(x, (tuple, list)):
        for i, v in enumerate(x):
            flat(v, path + "[%d]" % i, acc)
    return acc


h_inp, h_bat = bench.make_batch(4321, pin=False, batch=B)
inp = {k: v.to(dev) for k, v in h_inp.items()}
bat = {k: v.to(dev) for k, v in h_bat.items()}
tr = HotPathTrainer(dev, world_size=1, dropout=False)
tr.init_weights(seed=77)
tr.engine.use_streams = streams
st = tr.store
first = None
for value in (0.0, 3.0e4, float("nan"), -7.0):
    fill[0] = value
    st.grad.zero_()
    out, ctx = tr.engine.forward(inp, train=True)
    _, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, bat)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    fill[0] = None
    saved = [(p, t.clone()) for p, t in flat(ctx, "ctx", []) + flat(out, "out", []) + flat((d_hg, d_rel, d_act), "dlogit", [])]
    grad = st.grad.clone()
    if first is None:
        first = (saved, grad)
        continue
    print("=== fill %r vs fill 0" % value, flush=True)
    n_bad = 0
    for (p, a), (q, b) in zip(first[0], saved):
        if a.shape != b.shape or a.dtype != b.dtype or not a.is_floating_point():
            continue
        same = (a == b) | (torch.isnan(a) & torch.isnan(b))
        bad = int((~same).sum())
        if bad:
            n_bad += 1
            if n_bad <= 30:
                extra = ""
                if a.dim() == 2:
                    rows = torch.nonzero((~same).any(dim=1)).flatten()
                    cols = torch.nonzero((~same).any(dim=0)).flatten()
                    extra = " rows %d..%d (%d) cols %d..%d (%d)" % (int(rows[0]), int(rows[-1]), rows.numel(), int(cols[0]), int(cols[-1]), cols.numel())
                print("   saved %-36s %s %s differing %d / %d%s" % (p, tuple(a.shape), str(a.dtype)[6:], bad, a.numel(), extra), flush=True)
    print("   saved tensors differing: %d of %d" % (n_bad, len(saved)), flush=True)
    rows = []
    for k in sorted(st.live_keys, key=lambda k: st.offsets[k][0]):
        off, n, shape = st.offsets[k]
        a, b = first[1][off:off + n].double(), grad[off:off + n].double()
        r = float((a - b).norm() / a.norm().clamp_min(1e-300))
        if not (r <= 2e-6) and not k.endswith("self.key.bias"):
            rows.append((r, k, shape))
    print("   parameter gradients with rel-L2 > 2e-6 (or NaN): %d" % len(rows), flush=True)
    for r, k, shape in rows[:40]:
        print("      %.2e %s %s" % (r, k, tuple(shape)), flush=True)
