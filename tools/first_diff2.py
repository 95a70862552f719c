"""Non-intrusive variant of first_diff.py: free-running steps (no host synchronisation between launches); after each
step compare (1) every tensor the forward saved for the backward and (2) the gradient of every parameter, in flat
(= roughly execution) order, with the first step.   python tools/first_diff2.py [B] [streams] [runs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from shg_vqa_b200.trainer import HotPathTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
streams = len(sys.argv) > 2 and sys.argv[2] == "streams"
n_runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3


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
for it in range(n_runs):
    st.grad.zero_()
    out, ctx = tr.engine.forward(inp, train=True)
    _, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, bat)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    saved = [(p, t.clone()) for p, t in flat(ctx, "ctx", []) + flat(out, "out", []) + flat((d_hg, d_rel, d_act), "dlogit", [])]
    grad = st.grad.clone()
    if first is None:
        first = (saved, grad)
        continue
    print("=== run %d vs run 0" % it, flush=True)
    n_bad = 0
    for (p, a), (q, b) in zip(first[0], saved):
        if a.shape != b.shape or a.dtype != b.dtype:
            continue
        if a.is_floating_point():
            d = (a.float() - b.float())
            bad = int((d != 0).sum())
        else:
            bad = int((a != b).sum())
        if bad:
            n_bad += 1
            if n_bad <= 25:
                extra = ""
                if a.is_floating_point() and a.dim() == 2:
                    rows = torch.nonzero((d != 0).any(dim=1)).flatten()
                    cols = torch.nonzero((d != 0).any(dim=0)).flatten()
                    extra = " rows %d..%d (%d) cols %d..%d (%d) max|d| %.3e" % (int(rows[0]), int(rows[-1]), rows.numel(), int(cols[0]),
                                                                    int(cols[-1]), cols.numel(), float(d.abs().max()))
                print("   saved %-40s %s differing %d / %d%s" % (p, tuple(a.shape), bad, a.numel(), extra), flush=True)
    print("   saved tensors differing: %d of %d" % (n_bad, len(saved)), flush=True)
    rows = []
    for k in sorted(st.live_keys, key=lambda k: st.offsets[k][0]):
        off, n, shape = st.offsets[k]
        a, b = first[1][off:off + n].double(), grad[off:off + n].double()
        r = float((a - b).norm() / a.norm().clamp_min(1e-300))
        rows.append((r, k, shape))
    rows_bad = [x for x in rows if x[0] > 2e-6]
    print("   parameter gradients with rel-L2 > 2e-6: %d of %d" % (len(rows_bad), len(rows)), flush=True)
    for r, k, shape in rows_bad[:60]:
        print("      %.2e %s %s" % (r, k, tuple(shape)), flush=True)
