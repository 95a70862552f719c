"""One-GPU emulation of bench.dp_check: the `world` shards of 4 clips run one after the other through ONE trainer
(gradients accumulate in the flat buffer, the CE denominators are the precomputed global sums) and are compared
with the concatenated batch, per parameter group.  Tells a reduction problem (cannot show here) from a batch-size
dependent code path (shows here).   python tools/dp_emul.py [world ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from shg_vqa_b200.trainer import HotPathTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
worlds = [int(x) for x in sys.argv[1:]] or [2, 4, 8]
per = 4


def cat_batches(shards):
    cat_inp = {k: torch.cat([s[0][k] for s in shards]).to(dev) for k in shards[0][0]}
    cat_bat = {"target": torch.cat([s[1]["target"] for s in shards]).to(dev)}
    for fam in ("rel", "act"):
        labs = [s[1][fam + "_labels"] for s in shards]
        offs, base = [shards[0][1][fam + "_off"][:1]], 0
        for s_ in shards:
            o = s_[1][fam + "_off"]
            offs.append(o[1:] + base)
            base += int(o[-1])
        cat_bat[fam + "_labels"] = torch.cat(labs).to(dev)
        cat_bat[fam + "_off"] = torch.cat(offs).to(torch.int32).to(dev)
    return cat_inp, cat_bat


def grads_of(tr, inp, bat):
    out, ctx = tr.engine.forward(inp, train=True)
    scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, bat)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    return out, scal.clone()


for world in worlds:
    shards = [bench.make_batch(4321 + r, pin=False, batch=per) for r in range(world)]
    # single process, concatenated batch
    one = HotPathTrainer(dev, world_size=1, dropout=False)
    one.init_weights(seed=77)
    cat_inp, cat_bat = cat_batches(shards)
    out_cat, scal_cat = grads_of(one, cat_inp, cat_bat)
    g_one = one.store.grad[one.store.live_extent].clone()
    # sequential shards; pass 1 collects the global CE sums, pass 2 uses them
    seq = HotPathTrainer(dev, world_size=1, dropout=False)
    seq.init_weights(seed=77)
    den = torch.zeros(8, device=dev)
    outs = []
    for r in range(world):
        inp = {k: v.to(dev) for k, v in shards[r][0].items()}
        bat = {k: v.to(dev) for k, v in shards[r][1].items()}
        out, _ = seq.engine.forward(inp, train=True)
        scal, *_ = seq.loss_and_grads(out, bat, want_grad=False)
        den[0:4] += scal[1:5]
        den[4:8] += scal[5:9]
        outs.append({k: v.clone() for k, v in out.items() if torch.is_tensor(v)})
    seq.store.grad.zero_()
    seq.world_size, seq.global_ce_norm = world, True
    real_all_reduce = dist.all_reduce
    dist.all_reduce = lambda t, group=None: t.copy_(den)
    try:
        for r in range(world):
            inp = {k: v.to(dev) for k, v in shards[r][0].items()}
            bat = {k: v.to(dev) for k, v in shards[r][1].items()}
            grads_of(seq, inp, bat)
    finally:
        dist.all_reduce = real_all_reduce
    g_seq = seq.store.grad[seq.store.live_extent] / world
    rel = float((g_seq.double() - g_one.double()).norm() / g_one.double().norm())
    print("world %d: |g_seq - g_cat| / |g_cat| = %.3e   CE sums cat %s  seq %s" % (
        world, rel, [round(float(x), 4) for x in scal_cat[1:9]], [round(float(x), 4) for x in den]), flush=True)
    # forward outputs of the same clips: bit-identical?
    for k in outs[0]:
        a = torch.cat([o[k] for o in outs]).float()
        b = out_cat[k].float()
        if a.shape == b.shape:
            d = (a - b).abs()
            print("    fwd %-12s max diff %.3e  differing %d / %d" % (k, float(d.max()), int((d > 0).sum()), d.numel()), flush=True)
    st = one.store
    lb = st.live_begin
    for name, sls in st.dp_groups.items():
        num = den_ = 0.0
        for sl in sls:
            a = g_seq[sl.start - lb:sl.stop - lb].double()
            b = g_one[sl.start - lb:sl.stop - lb].double()
            num += float((a - b).pow(2).sum()); den_ += float(b.pow(2).sum())
        print("    group %-6s rel-L2 %.3e" % (name, (num / max(den_, 1e-300)) ** 0.5), flush=True)
    del one, seq
    torch.cuda.empty_cache()
