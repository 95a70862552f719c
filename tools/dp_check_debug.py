"""Where does the data-parallel gradient differ from the single-process one?  (torchrun, N >= 2 GPUs.)
Per rank: local gradient of its 4-clip shard (global CE sums via NCCL).  Rank 0 then (a) recomputes every shard's
gradient itself with the same global sums and compares with the gathered one, (b) compares NCCL's SUM with the fp64
sum of the gathered gradients, (c) compares with the concatenated batch, per parameter group.
python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/dp_check_debug.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from shg_vqa_b200.trainer import HotPathTrainer

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = 4
shards = [bench.make_batch(4321 + r, pin=False, batch=per) for r in range(world)]


def to_dev(s):
    return {k: v.to(dev) for k, v in s[0].items()}, {k: v.to(dev) for k, v in s[1].items()}


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def groups(st, a, b, tag):
    lb = st.live_begin
    for name, sls in st.dp_groups.items():
        num = den = 0.0
        for sl in sls:
            x, y = a[sl.start - lb:sl.stop - lb].double(), b[sl.start - lb:sl.stop - lb].double()
            num += float((x - y).pow(2).sum()); den += float(y.pow(2).sum())
        print("    %s group %-6s rel-L2 %.3e" % (tag, name, (num / max(den, 1e-300)) ** 0.5), flush=True)


chk = HotPathTrainer(dev, world_size=world, dropout=False, global_ce_norm=True, rank=rank)
chk.init_weights(seed=77)
inp, bat = to_dev(shards[rank])
out, ctx = chk.engine.forward(inp, train=True)
_, d_hg, d_rel, d_act, _ = chk.loss_and_grads(out, bat)
den = chk.scal_den.clone()
chk.engine.backward(ctx, d_hg, d_rel, d_act)
torch.cuda.synchronize()
live = chk.store.live_extent
g_local = chk.store.grad[live].clone()
gathered = [torch.empty_like(g_local) for _ in range(world)] if rank == 0 else None
dist.gather(g_local, gathered, dst=0)
chk.all_reduce_grads()
torch.cuda.synchronize()
g_nccl = chk.store.grad[live].clone()
dens = [torch.empty_like(den) for _ in range(world)]
dist.all_gather(dens, den)
if rank == 0:
    print("global CE sums per rank identical:", all(torch.equal(dens[0], d) for d in dens), dens[0].tolist(), flush=True)
    acc = torch.zeros_like(g_local, dtype=torch.float64)
    for g in gathered:
        acc += g.double()
    print("(b) NCCL sum vs fp64 sum of the gathered gradients: rel-L2 %.3e" % rel_l2(g_nccl, acc), flush=True)
    # (a) recompute every shard here
    loc = HotPathTrainer(dev, world_size=1, dropout=False)
    loc.init_weights(seed=77)
    loc.world_size, loc.global_ce_norm = world, True
    real = dist.all_reduce
    dist.all_reduce = lambda t, group=None: t.copy_(dens[0])
    for r in range(world):
        loc.store.grad.zero_()
        i2, b2 = to_dev(shards[r])
        o2, c2 = loc.engine.forward(i2, train=True)
        _, a1, a2, a3, _ = loc.loss_and_grads(o2, b2)
        loc.engine.backward(c2, a1, a2, a3)
        torch.cuda.synchronize()
        g = loc.store.grad[live]
        print("(a) shard %d: gradient computed on rank %d vs recomputed on rank 0: rel-L2 %.3e" % (r, r, rel_l2(gathered[r], g)), flush=True)
        if rel_l2(gathered[r], g) > 1e-5:
            groups(loc.store, gathered[r], g, "shard %d" % r)
    dist.all_reduce = real
    del loc
    one = HotPathTrainer(dev, world_size=1, dropout=False)
    one.init_weights(seed=77)
    cat_inp = {k: torch.cat([s[0][k] for s in shards]).to(dev) for k in shards[0][0]}
    cat_bat = {"target": torch.cat([s[1]["target"] for s in shards]).to(dev)}
    for fam in ("rel", "act"):
        offs, base = [shards[0][1][fam + "_off"][:1]], 0
        for s_ in shards:
            o = s_[1][fam + "_off"]
            offs.append(o[1:] + base)
            base += int(o[-1])
        cat_bat[fam + "_labels"] = torch.cat([s[1][fam + "_labels"] for s in shards]).to(dev)
        cat_bat[fam + "_off"] = torch.cat(offs).to(torch.int32).to(dev)
    o3, c3 = one.engine.forward(cat_inp, train=True)
    s3, a1, a2, a3, _ = one.loss_and_grads(o3, cat_bat)
    one.engine.backward(c3, a1, a2, a3)
    torch.cuda.synchronize()
    g_one = one.store.grad[live]
    print("concatenated CE sums:", s3[1:9].tolist(), flush=True)
    print("(c) NCCL sum / world vs concatenated batch: rel-L2 %.3e" % rel_l2(g_nccl / world, g_one), flush=True)
    groups(one.store, g_nccl / world, g_one, "dp vs cat")
    print("(c') fp64 gathered sum / world vs concatenated batch: rel-L2 %.3e" % rel_l2(acc / world, g_one), flush=True)
dist.barrier()
dist.destroy_process_group()
