"""Which launch makes two runs of the same step differ?  One GPU, every launch on one stream.  Wraps the host-side
ops, fingerprints every tensor an op returns / writes, runs the step twice and lists the first calls whose
fingerprints differ.      python tools/first_diff.py [B] [streams]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shg_vqa_b200 import ops
from shg_vqa_b200.trainer import HotPathTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
streams = len(sys.argv) > 2 and sys.argv[2] == "streams"
log = []


def fp(t):
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    bits = t.view(torch.int16 if t.element_size() == 2 else torch.int32).to(torch.int64)
    return (int(bits.sum()), int((bits * (torch.arange(bits.numel(), device=t.device).view(bits.shape) % 8191 + 1)).sum()))


def tensors_of(x):
    if torch.is_tensor(x):
        return [x]
    if isinstance(x, (tuple, list)):
        return [t for y in x for t in tensors_of(y)]
    return []


def wrap(name, extra=()):
    real = getattr(ops, name)

    def f(*a, **kw):
        res = real(*a, **kw)
        outs = tensors_of(res) + [kw[k] for k in ("out",) if torch.is_tensor(kw.get(k))] + [a[i] for i in extra if i < len(a)]
        torch.cuda.synchronize()
        desc = "%s %s" % (name, " ".join("%s%s" % (k, tuple(v.shape)) for k, v in
                                         list(zip("abcdefghij", [x for x in a if torch.is_tensor(x)][:3]))))
        flags = ",".join(k for k, v in kw.items() if v is not None and v is not False and k not in ("out",))
        log.append((desc + " [" + flags + "]", [fp(t) for t in outs if t is not None and t.dtype in (torch.bfloat16, torch.float32, torch.int32)]))
        return res
    setattr(ops, name, f)


for n, extra in (("gemm", ()), ("attention_fwd", ()), ("attention_bwd", (6, 7, 8)), ("layernorm_fwd", ()), ("layernorm_bwd", ()),
                 ("add_rows", (0,)), ("act_bwd", ()), ("bce_logits", ()), ("copy3d", (1,)), ("set_loss_bwd", ()),
                 ("match_from_logits", ())):
    wrap(n, extra)

h_inp, h_bat = bench.make_batch(4321, pin=False, batch=B)
inp = {k: v.to(dev) for k, v in h_inp.items()}
bat = {k: v.to(dev) for k, v in h_bat.items()}
tr = HotPathTrainer(dev, world_size=1, dropout=False)
tr.init_weights(seed=77)
tr.engine.use_streams = streams
runs = []
for it in range(3):
    log.clear()
    tr.store.grad.zero_()
    out, ctx = tr.engine.forward(inp, train=True)
    n_fwd = len(log)
    _, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, bat)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    runs.append((list(log), n_fwd, tr.store.grad[tr.store.live_extent].clone()))
    del out, ctx
for it in (1, 2):
    a, b = runs[0], runs[it]
    rel = float((a[2].double() - b[2].double()).norm() / a[2].double().norm())
    print("run %d vs run 0: gradient rel-L2 %.3e, %d calls (%d forward)" % (it, rel, len(b[0]), b[1]), flush=True)
    shown = 0
    for i, (x, y) in enumerate(zip(a[0], b[0])):
        if x[1] != y[1]:
            print("   call %4d %s  %s" % (i, "fwd" if i < b[1] else "bwd", x[0]), flush=True)
            shown += 1
            if shown >= 12:
                break
