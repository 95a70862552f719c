"""Data-parallel divergence probe (run under torchrun, 2+ GPUs): after N steps, which part of the live span differs
between ranks?  python -m torch.distributed.run --nproc-per-node 2 tools/dp_debug.py [eager|graph] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from shg_vqa_b200.trainer import HotPathTrainer

mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tr = HotPathTrainer(dev, world_size=world, t_total=100000, warmup=0.1, lr=1e-5, dropout=True, dropout_seed=1, rank=rank)
tr.init_weights(seed=9595)
h_inp, h_batch = bench.make_batch(1234 + rank, pin=False)
inp = {k: v.to(dev) for k, v in h_inp.items()}
batch = {k: v.to(dev) for k, v in h_batch.items()}


def report(tag, t):
    st = tr.store
    other = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(other, t.contiguous())
    if rank == 0:
        d = (other[0] - other[1]).abs()
        print("%s: max |rank0 - rank1| = %.3e  (differing elements %d of %d)" % (tag, float(d.max()), int((d > 0).sum()), d.numel()), flush=True)
        names = [("late", [st.bucket_late]), ("upper", st.bucket_early_upper), ("rest", st.bucket_early_rest)]
        for n, sls in names:
            tot = bad = 0
            for sl in sls:
                seg = d[sl.start - st.live_begin:sl.stop - st.live_begin]
                tot += seg.numel(); bad += int((seg > 0).sum())
            print("    bucket %-6s differing %d / %d" % (n, bad, tot), flush=True)
        if int((d > 0).sum()):
            idx = torch.nonzero(d > 0).flatten()[:5] + st.live_begin
            for i in idx.tolist():
                k = max((kk for kk in st.offsets if st.offsets[kk][0] <= i), key=lambda kk: st.offsets[kk][0])
                print("    e.g. flat %d in %s" % (i, k), flush=True)


live = tr.store.live_extent
report("initial weights", tr.store.master[live])
if mode == "eager":
    for _ in range(steps):
        tr.train_step(inp, batch)
elif mode == "gradonly":
    # one eager backward + reduce, no optimiser: are the REDUCED gradients identical?
    tr._advance_seed()
    out, ctx = tr.engine.forward(inp, train=True)
    scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, batch)
    st = tr.store
    tr.engine.backward_late_params(ctx, d_hg, d_rel, d_act)
    works = tr._reduce_async([st.bucket_late])
    tr.engine.backward_early_params(ctx, part=0, join=False)
    works += tr._reduce_async(st.bucket_early_upper)
    tr.engine.backward_early_params(ctx, part=1)
    works += tr._reduce_async(st.bucket_early_rest)
    for w in works:
        w.wait()
    torch.cuda.synchronize()
    report("reduced gradients (overlapped schedule)", tr.store.grad[live])
    sys.exit(0)
else:
    tr.capture(inp, batch)
    report("after capture (state restored)", tr.store.master[live])
    for _ in range(steps):
        tr.train_step_graph()
torch.cuda.synchronize()
report("weights after %d %s steps" % (steps, mode), tr.store.master[live])
report("exp_avg", tr.exp_avg[live])
tr.release()
dist.destroy_process_group()
